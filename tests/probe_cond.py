"""Timing of the conditioning path: kernels (csrc/cond.cu) vs the torch statement, forward and forward+backward, CUDA events."""
import sys
import torch
from videometamaterials_b200 import Unet3D, blocks
from videometamaterials_b200.blocks_bwd import get_arena

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
               cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True).cuda()
arena = get_arena(model)
time = torch.randint(0, 256, (B,), device="cuda")
cond = torch.rand(B, 11, device="cuda") * 2 - 1
null = torch.rand(B, device="cuda") < 0.25


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def k_fwd():
    return blocks.conditioning_state(model, time, cond, null, 11)


def k_fwd_bwd():
    *_, st = k_fwd()
    st.grad_views()
    st.backward()


def t_fwd():
    return blocks.conditioning(model, time, cond, null, 11, _torch_path=True)


def t_fwd_bwd():
    ss, ekv, bias, rot = t_fwd()
    (sum(v.sum() for v in ss.values()) + sum(v.sum() for v in ekv.values()) + bias.sum()).backward()


print(f"B={B}: kernels fwd {timed(k_fwd):.0f} us, fwd+bwd {timed(k_fwd_bwd):.0f} us | torch fwd {timed(t_fwd):.0f} us, fwd+bwd {timed(t_fwd_bwd):.0f} us (eager launches, host-bound)")
# device time of the kernels alone: the same parameter block launched back to back (no Python plan / allocation in the loop)
from videometamaterials_b200 import ops
*_, st = k_fwd()
st.grad_views()
p = st.p
fwd = lambda: ops.check(ops.lib.vmm_cond_fwd(ops.C.byref(p), ops.stream_ptr()))
print(f"  vmm_cond_fwd device time {timed(fwd, 50):.0f} us")
p.out = st.dout.data_ptr()
p.grad = arena.flat_grad.data_ptr()
bwd = lambda: ops.check(ops.lib.vmm_cond_bwd(ops.C.byref(p), ops.stream_ptr()))
print(f"  vmm_cond_bwd device time {timed(bwd, 50):.0f} us (memset + main + tail)")
