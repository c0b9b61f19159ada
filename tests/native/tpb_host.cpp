// Test-only host build of csrc/xq_rules_tpb.h (the scalar rules the thread-per-board kernel runs):
// g++ compiles the very same header, tests/test_tpb_cpu.py compares it with the oracle.  Not part of
// libxq_b200.so -- the product has no CPU path.
#include <cstdint>
#include <cstring>
#include "../../xiangqi-alphazero_b200/csrc/xq_rules_tpb.h"

extern "C" int xqt_host_movegen_batch(const int8_t* boards, const int8_t* sides, int B, int16_t* actions,
                                      uint8_t* n_moves, uint8_t* in_check)
{
    int overflow = 0;
    uint32_t tab[xqt::kSlotTableSize];
    for (int i = 0; i < xqt::kSlotTableSize; ++i) tab[i] = xqt::slot_entry(i);
    for (int i = 0; i < B; ++i) {
        int8_t padded[32 + 90 + 32] = {0};
        int8_t* b = padded + 32;                  // the generator may read (never write) up to 20 bytes outside the board
        std::memcpy(b, boards + (size_t)i * 90, 90);
        uint16_t list[xqt::kListCap];
        int16_t out[128];
        int chk = 0;
        uint32_t occ[3];
        int nf = 0;
        int n = xqt::movegen(b, sides[i], list, out, &nf, &chk, tab, occ);
        for (int k = nf; k < n && k < 128; ++k) out[k] = (int16_t)list[k - nf];
        for (int q = 0; q < 90; ++q)
            if (((occ[q >> 5] >> (q & 31)) & 1u) != (b[q] != 0 ? 1u : 0u)) return -1 - i;
        if (std::memcmp(b, boards + (size_t)i * 90, 90) != 0) return -1 - i;   // the board must come back untouched
        for (int q = 0; q < 32; ++q)
            if (padded[q] != 0 || padded[32 + 90 + q] != 0) return -1 - i;         // ... and nothing written around it
        if (n > 128) { ++overflow; n = 128; }
        for (int k = 0; k < 128; ++k) actions[(size_t)i * 128 + k] = k < n ? out[k] : (int16_t)-1;
        n_moves[i] = (uint8_t)n;
        in_check[i] = (uint8_t)chk;
    }
    return overflow;
}
