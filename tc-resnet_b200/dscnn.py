"""DS-CNN forward (inference) on the CUDA path: the reference's 2-D-conv comparison model
(audio_nets/ds_cnn.py; DSCNN{S,M}Model.build_inference in factory/audio_nets.py:298-358).  Training DS-CNN is out of scope
(SURVEY.md 2 #13): weights come from a checkpoint keyed by the TF variable names in `self.table`."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import _lib as L


class DsCnn:
    def __init__(self, size: str = "S", height: int = 49, width: int = 40, num_classes: int = 12, max_batch: int = 512,
                 device: int | None = None):
        if not torch.cuda.is_available():
            raise L.TcrError("tcresnet_b200.DsCnn needs a CUDA device (there is no CPU fallback)")
        self.lib = L.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        cfg = L.TcrDscnnConfig(ord(size), height, width, num_classes, max_batch, self.device.index)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib, self.lib.tcr_dscnn_create(C.byref(cfg), C.byref(self._h)), "tcr_dscnn_create")
        descs, count, nparams, flops = C.POINTER(L.TcrParamDesc)(), C.c_int32(), C.c_int64(), C.c_int64()
        L.check(self.lib, self.lib.tcr_dscnn_param_table(self._h, C.byref(descs), C.byref(count), C.byref(nparams),
                                                         C.byref(flops)), "tcr_dscnn_param_table")
        self.table = [dict(name=descs[i].name.decode(), shape=tuple(descs[i].shape[:descs[i].rank]), offset=int(descs[i].offset),
                           numel=int(descs[i].numel)) for i in range(count.value)]
        self.num_params, self.forward_flops = int(nparams.value), int(flops.value)
        self.num_classes, self.height, self.width = num_classes, height, width

    def close(self):
        if getattr(self, "_h", None):
            self.lib.tcr_dscnn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pack(self, values: Dict[str, np.ndarray]) -> torch.Tensor:
        flat = np.zeros(self.num_params, np.float32)
        for d in self.table:
            flat[d["offset"]:d["offset"] + d["numel"]] = np.asarray(values[d["name"]], np.float32).ravel()
        return torch.from_numpy(flat).to(self.device)

    def forward(self, features: torch.Tensor, params: torch.Tensor):
        n = features.shape[0]
        features = features.reshape(n, self.height, self.width).contiguous()
        logits = torch.empty(n, self.num_classes, dtype=torch.float32, device=self.device)
        probs = torch.empty_like(logits)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib, self.lib.tcr_dscnn_forward(self._h, features.data_ptr(), params.data_ptr(), n, logits.data_ptr(),
                                                     probs.data_ptr(), stream), "tcr_dscnn_forward")
        return logits, probs
