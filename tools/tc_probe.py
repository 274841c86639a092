"""First-contact check of the tensor-core layer-1 kernel (csrc/rb_head_tc.cu): run under `timeout` on a GPU box.

Compares rb_head_forward with the tcgen05 kernel against (a) the FFMA kernel and (b) composed weights + torch fp32 matmul, for
the learner's shapes, and times both layer-1 implementations with CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rainbow_b200 import _lib  # noqa: E402
from test_gpu_parity import _head_net, _library_head  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
L = _lib.load()
for arch, hidden, actions, rows in (("canonical", 512, 6, (32, 32)), ("canonical", 512, 6, (32, 0)), ("data-efficient", 256, 6, (16, 16)),
                                    ("canonical", 512, 6, (1, 0)), ("data-efficient", 64, 3, (40, 60))):
    net = _head_net(arch, hidden, actions)
    net.reset_noise()
    torch.manual_seed(1)
    m_lo, m_hi = rows
    x_lo = torch.relu(torch.randn(m_lo, net.conv_output_size, device="cuda"))
    x_hi = torch.relu(torch.randn(m_hi, net.conv_output_size, device="cuda")) if m_hi else None
    head = net.head()
    L.rb_head_debug(0)
    z_tc, h_tc, _ = head.forward(x_lo, x_hi)
    z_tc, h_tc = z_tc.clone(), h_tc.clone()
    torch.cuda.synchronize()
    L.rb_head_debug(4)
    z_ff, h_ff, _ = head.forward(x_lo, x_hi)
    z_ff, h_ff = z_ff.clone(), h_ff.clone()
    torch.cuda.synchronize()
    L.rb_head_debug(0)
    feats = x_lo if x_hi is None else torch.cat([x_lo, x_hi])
    import torch.nn.functional as F
    with torch.no_grad():
        v, a = _library_head(net, feats)
        z_ref = torch.cat([v, a], 1)
        h_ref = torch.cat([F.relu(net.fc_h_v(feats)), F.relu(net.fc_h_a(feats))], 1)
    print(f"{arch}/{hidden} rows {rows}: |h_tc-h_ff| {float((h_tc - h_ff).abs().max()):.3e}  |h_tc-h_ref| {float((h_tc - h_ref).abs().max()):.3e}  "
          f"|h_ff-h_ref| {float((h_ff - h_ref).abs().max()):.3e}  |z_tc-z_ref| {float((z_tc - z_ref).abs().max()):.3e}  scale {float(h_ref.abs().max()):.2f}", flush=True)
    for flag, name in ((0, "tcgen05"), (4, "ffma")):
        L.rb_head_debug(flag | 2)                 # layer 1 only
        for _ in range(5):
            head.forward(x_lo, x_hi)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            head.forward(x_lo, x_hi)
        e1.record()
        torch.cuda.synchronize()
        print(f"    layer 1 {name}: {e0.elapsed_time(e1) * 1e3 / 50:.2f} us per call (back to back, L2-warm)", flush=True)
    L.rb_head_debug(0)
print("tc probe done")
