"""Runs a GPU-only script of this repo (bench.py, tools/mgpu_check.py, ...) without a GPU -- TEST INFRASTRUCTURE
(tests/test_bench_dry_run.py):      python tests/emul/run_on_executor.py bench.py --particles 12000 --steps 2 ...

Those scripts can only run on a CUDA device; a Python-level mistake in them would otherwise first show up on the GPU box at
round end.  This launcher makes the same code run here: the library handle is the CPU executor of the CUDA sources
(tests/emul/cuda_emul.h), torch's CUDA entry points used by bench.py / distributed.Runner are mapped to host equivalents
(events -> wall clock, pinned / device tensors -> plain host tensors), NCCL -> gloo.  The numbers such a run prints are
meaningless; only the control flow and the shape of the output are checked."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


class _Event:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


def _cpu_device(kwargs):
    if "device" in kwargs and str(kwargs["device"]).startswith("cuda"):
        kwargs["device"] = "cpu"
    return kwargs


torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *_a, **_k: None
torch.cuda.synchronize = lambda *_a, **_k: None
torch.cuda.Event = _Event
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.Tensor.cuda = lambda self, *a, **k: self
_tensor, _empty, _full, _zeros = torch.tensor, torch.empty, torch.full, torch.zeros
_cpu = torch.Tensor.cpu
torch.tensor = lambda *a, **k: _tensor(*a, **_cpu_device(k))
torch.empty = lambda *a, **k: _empty(*a, **_cpu_device(k))
torch.full = lambda *a, **k: _full(*a, **_cpu_device(k))
torch.zeros = lambda *a, **k: _zeros(*a, **_cpu_device(k))
_init = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _init("gloo", **{q: v for q, v in k.items() if q != "device_id"})

import splashsurf_b200 as ss  # noqa: E402
from splashsurf_b200 import distributed as ssd  # noqa: E402
from test_emulated_pipeline import build_emulated_library  # noqa: E402

ss._LIB = ss._bind(ctypes.CDLL(build_emulated_library()))
_Runner = ssd.Runner
ssd.Runner = lambda *a, **k: _Runner(*a, **dict(k, device="cpu"))

if __name__ == "__main__":
    import runpy
    script = sys.argv[1] if os.path.isabs(sys.argv[1]) else os.path.join(ROOT, sys.argv[1])
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")
