"""Find the first decode step where the CUDA guided beam search and the oracle disagree."""
import sys, torch
sys.path.insert(0, ".")
from oracle import spectra_oracle as orc
from oracle.guided_oracle import GuidedOracle
from tests.helpers import load_case, oracle_cfg
from tests.toy_chem import ToyChem
from tests.test_guided import VocabTokenizer, golden, _build
from multimodalanalytical_b200.guided import GuidedFormulaProcessor

K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = golden()
fx, tok, m = _build()
cfg = oracle_cfg(fx)
trace_o = []
class Rec(GuidedOracle):
    def __call__(self, ids, scores):
        trace_o.append((ids.clone(), torch.from_numpy(self.counts(ids)).clone()))
        return super().__call__(ids, scores)
want = orc.generate(fx["state_dict"], cfg, fx["batch"], n_beams=K, logits_hook=[Rec(K, g["formulas"], fx["smiles_vocab"], 3, ToyChem())])
trace_g = []
class RecG(GuidedFormulaProcessor):
    def counts(self, ids, out=None):
        r = super().counts(ids, out)
        trace_g.append((ids.clone().long(), r.clone()))
        return r
got = m.generate(fx["batch"], n_beams=K, logits_processor=[RecG(K, g["formulas"], tok, chem=ToyChem())]).cpu()
print("steps oracle", len(trace_o), "gpu", len(trace_g), "shapes", want.shape, got.shape)
n = min(want.shape[1], got.shape[1])
bad = sorted({r // K for r in range(want.shape[0]) if not torch.equal(want[r, :n], got[r, :n]) or (want[r, n:] != 3).any()})
print("spectra with different final hypotheses:", bad)
for b in bad[:2]:
    print("spectrum", b, "target", g["formulas"][b])
    print(" want", [tok.batch_decode(want[b*K:(b+1)*K])[i].replace(" ", "") for i in range(K)])
    print(" got ", [tok.batch_decode(got[b*K:(b+1)*K])[i].replace(" ", "") for i in range(K)])
    for t, ((io, co), (ig, cg)) in enumerate(zip(trace_o, trace_g)):
        if not torch.equal(io[b*K:(b+1)*K], ig[b*K:(b+1)*K]):
            print(" first differing step", t + 1)
            print("  oracle", io[b * K:(b + 1) * K].tolist())
            print("  gpu   ", ig[b * K:(b + 1) * K].tolist())
            print("  prev counts oracle", trace_o[t - 1][1][b * K:(b + 1) * K].long().tolist())
            print("  prev counts gpu   ", trace_g[t - 1][1][b * K:(b + 1) * K].tolist())
            break
