"""Compatibility shim for the reference's training/inference_server.py.

The reference runs one spawned server process that batches single-position requests arriving over
Unix sockets (inference_server.py:37-297).  Here batching happens on the device (all leaves of a
search step are one forward, see selfplay_engine), so no server process exists; this module keeps
the class names and call shapes for callers that still construct them:

    InferenceServer(model_class, model_kwargs, state_dict, device, num_workers, max_batch_size,
                    batch_timeout_ms).start() -> str ; .stop() ; .update_model(state_dict)
    InferenceClient(worker_id, socket_path).predict(state) -> (float32[8100], float) ; .close()
"""
import numpy as np

_MODELS = {}


class InferenceServer:
    def __init__(self, model_class, model_kwargs: dict, state_dict: dict, device: str = 'cuda', num_workers: int = 4,
                 max_batch_size: int = 32, batch_timeout_ms: float = 5.0):
        self.model = model_class(**model_kwargs)
        self.model.load_state_dict(state_dict)
        self.model.eval()
        self.socket_path = None

    def start(self) -> str:
        self.socket_path = f"xq-b200-inproc-{id(self):x}"
        _MODELS[self.socket_path] = self.model
        return self.socket_path

    def stop(self):
        _MODELS.pop(self.socket_path, None)

    def update_model(self, state_dict: dict):
        self.model.load_state_dict(state_dict)      # next predict() refolds the weights (model.b200())

    def __del__(self):
        try:
            self.stop()
        except Exception:
            pass


class InferenceClient:
    def __init__(self, worker_id: int, socket_path: str):
        self.worker_id = worker_id
        self.socket_path = socket_path

    def predict(self, state: np.ndarray):
        model = _MODELS.get(self.socket_path)
        if model is None:
            raise RuntimeError("inference server is not running")       # inference_server.py:348-349
        return model.predict(state)

    def close(self):
        pass
