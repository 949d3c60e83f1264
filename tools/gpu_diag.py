import ctypes as C
import torch
from videometamaterials_b200 import _lib
torch.zeros(1).cuda()
out = (C.c_int * 12)()
print("rc", _lib.lib.vmm_ftattn_diag(out))
names = ["numRegs", "static_smem", "max_dyn_smem", "smem_per_sm", "reserved_per_block", "regs_per_sm", "occ@FT_SMEM", "occ@114944", "occ@113664", "occ@106496", "occ@65536", "local_bytes"]
print({n: out[i] for i, n in enumerate(names)})
