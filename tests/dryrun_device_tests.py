"""Pre-flight for the `-m gpu` tests on a machine without a GPU  --  TEST TOOLING, not collected by pytest.

    python tests/dryrun_device_tests.py

Runs the bodies of the whole-network device tests (and `__graft_entry__.smoke()`) with every kernel wrapper replaced by the torch
statement of its contract (tests/emu_ops.py, tests/emu_cgemm.py), `.cuda()` turned into a no-op that marks tensors as device
tensors, and 16-bit storage kept as the tests request it.  It catches Python-level breakage of the glue and of the tests themselves
before any GPU minute is spent, and its error figures track the device's closely (round 1: smoke forward 1.54e-3 here vs 1.53e-3 on
the B200; fp16 gradient-norm worst case 3.1 % vs 3.0 %).  It cannot say anything about the kernels, CUDA graphs or streams.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import emu_ops  # noqa: E402
from videometamaterials_b200 import _lib, ops  # noqa: E402


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


class _ClaimsCuda(torch.Tensor):
    is_cuda = property(lambda self: True)


def main():
    emu_ops.install_training(_Patch(), ops)
    emu_ops.install_sampler(_Patch(), ops)
    torch.Tensor.cuda = lambda self, *a, **k: (self.as_subclass(_ClaimsCuda) if self.is_floating_point() else self)
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *a, **k: None
    count = [0]

    def launches():
        count[0] += 500
        return count[0]
    _lib.launch_count = launches

    import __graft_entry__
    import test_gpu_unet as T
    gold = torch.load(os.path.join(HERE, "golden", "small_unet.pt"))
    steps = [("smoke()", __graft_entry__.smoke)]
    steps += [(f"small forward {dt}", lambda dt=dt: T.test_small_unet_forward(gold, dt)) for dt in (torch.float16, torch.bfloat16)]
    steps += [("small sampling (eager)", lambda: T.test_small_sampling(gold, torch.float16))]
    steps += [(f"small training {dt}", lambda dt=dt, sc=sc: T.test_small_training_loss_and_gradients(gold, dt, sc))
              for dt, sc in ((torch.bfloat16, 1.0), (torch.float16, 4096.0))]
    steps += [(f"ragged {a}", lambda a=a: T.test_ragged_sizes_against_the_oracle(*a)) for a in ((20, 3, True), (24, 1, True))]
    steps += [(f"wrap-mode network {m}", lambda m=m: T.test_wrap_mode_network_against_the_oracle(m)) for m in ("circular", "circular_1d")]
    steps += [(f"config flags {a}", lambda a=a: T.test_config_flags_network_against_the_oracle(*a))
              for a in ((False, "add", 16), (True, "concat", 16), (False, "concat", 64))]
    failed = 0
    for name, fn in steps:
        try:
            fn()
            print(f"[ok]   {name}")
        except Exception as e:  # noqa: BLE001
            failed += 1
            print(f"[FAIL] {name}: {type(e).__name__}: {str(e)[:300]}")
    print(f"{len(steps) - failed} of {len(steps)} device-test bodies pass under the contract statements")
    return 1 if failed else 0


if __name__ == "__main__":
    raise SystemExit(main())
