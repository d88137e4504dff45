import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import load_case
from tests.test_model_gpu import build
from tests.test_parity_r2_gpu import _teacher_forced_fp32_logits
fx = load_case("c1_ir_tiny")
m32, m16 = build(fx, "fp32"), build(fx, "bf16")
m32.eval(); m16.eval()
for ug in (True, False):
    g16 = m16.generate(fx["batch"], n_beams=1, use_graph=ug).cpu()
    g32 = m32.generate(fx["batch"], n_beams=1, use_graph=ug).cpu()
    print("graph", ug, "shapes", g16.shape, g32.shape)
    lg32 = _teacher_forced_fp32_logits(m32, fx["batch"], g16, 1)
    lg16 = _teacher_forced_fp32_logits(m16, fx["batch"], g16, 1)
    for r in range(min(4, g16.shape[0])):
        print("row", r, "bf16", g16[r, :24].tolist())
        print("       fp32", g32[r, :24].tolist())
        for i in range(g16.shape[1] - 1):
            tok = int(g16[r, i + 1])
            a32, a16 = int(lg32[r, i].argmax()), int(lg16[r, i].argmax())
            if tok != a32 or tok != a16:
                print(f"   pos {i}: decoded {tok}, tf-fp32 argmax {a32} (gap {float(lg32[r,i].max()-lg32[r,i,tok]):.3f}), tf-bf16 argmax {a16} (gap {float(lg16[r,i].max()-lg16[r,i,tok]):.3f}) maxabs {float(lg32[r,i].abs().max()):.2f}")
                break
            if tok == 3:
                break
