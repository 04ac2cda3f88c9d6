"""Occupancy facts of the co-scheduled kernel on this device for a range of dynamic shared-memory sizes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_attention_b200 import _lib
lib = _lib.load()
torch.cuda.init(); torch.zeros(1, device="cuda")
for smem in (60000, 100000, 110000, 114688, 115712):
    out = (ctypes.c_int * 6)()
    rc = lib.pa_debug_cosched_occupancy(smem, out)
    print(smem, rc, list(out), lib.pa_last_error())
