"""Shared helpers of the parity tests: backend selection (real CUDA library on a GPU box, CPU emulation of the
same kernels otherwise), seeded inputs, error logging."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HAS_CUDA = torch.cuda.is_available()
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libfreehunch_emu.so")


def use_emulation():
    """CPU box only: point the package's binding at the g++ build of the SAME kernel sources (tests/emu/build_emu.sh,
    tools/cuda_emu.h) so their index logic and algebra can be checked without a GPU.  This lives in the test harness:
    the product loader knows nothing about it, and the -m gpu tests assert that the real library is the one loaded."""
    assert not HAS_CUDA, "the emulation build is never used on a GPU box"
    if not os.path.exists(EMU_LIB):
        import subprocess
        subprocess.check_call(["bash", os.path.join(ROOT, "tests", "emu", "build_emu.sh")])
    from free_hunch_b200 import _lib, engine
    _lib._lib = _lib._bind(EMU_LIB)
    engine._check_device = lambda device: None
    return engine

DEVICE = "cuda" if HAS_CUDA else "cpu"
SUB = 97


def sub(a):
    if torch.is_tensor(a):
        a = a.detach().cpu().numpy()
    return np.asarray(a, np.float64).reshape(-1)[::SUB]


def make_inputs(seed, shape):
    return np.random.RandomState(seed).standard_normal(shape)


def rel_l2(a, b):
    if torch.is_tensor(a):
        a = a.detach().cpu().numpy()
    if torch.is_tensor(b):
        b = b.detach().cpu().numpy()
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


_LOG = os.path.join(ROOT, "gpurun_out", "parity_log.jsonl")


def record(name, **vals):
    """Append measured errors to gpurun_out/parity_log.jsonl (scratch; summarised into profiles/ by hand)."""
    try:
        os.makedirs(os.path.dirname(_LOG), exist_ok=True)
        with open(_LOG, "a") as f:
            f.write(json.dumps({"test": name, "device": DEVICE, **{k: (float(v) if isinstance(v, (int, float, np.floating)) else v) for k, v in vals.items()}}) + "\n")
    except OSError:
        pass
