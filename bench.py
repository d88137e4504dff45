#!/usr/bin/env python
"""Benchmark of the spectra -> SMILES hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

Headline metric (BASELINE.json, configs[1] = "C2"): training spectra/s of the IR-only structure-elucidation
model (custom_model.yaml: d=512, 6+6 layers, 8 heads, ffn 2048; Formula[B,15] + IR patches [B,21,75] -> S=36;
T=64 target tokens, V=200; bf16; per-GPU batch 256; AdamW + OneCycle + clip 1.0; dropout 0.1), one step =
forward + backward + optimiser on one synthetic batch.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 3247
PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
SWEEP_B, SWEEP_K = "1,4,8,32,128,256,512,1024", "10,30"
C2 = dict(B=256, S_formula=15, P=21, ps=75, T=64, V=200, d=512, layers=6, heads=8, ffn=2048)


def flops_per_sample_train(c, gated=False):
    """algorithmic forward MACs x 2 x 3 (fwd + bwd)."""
    d, f, S, T, V = c["d"], c["ffn"], c["S_formula"] + c["P"], c["T"], c["V"]
    return flops_train(S, T, V, d, f, c["layers"], c["layers"], c["P"] * c["ps"] * d, gated)


def flops_train(S, T, V, d, f, Le, Ld, emb_macs, gated):
    g = 1 if gated else 0
    enc = Le * (S * (4 * d * d + (2 + g) * d * f) + 2 * S * S * d)
    dec = Ld * (T * (6 * d * d + (2 + g) * d * f) + 2 * S * d * d + T * (T + 1) * d + 2 * T * S * d)
    return 2.0 * 3.0 * (enc + dec + emb_macs + T * d * V)


def _tokcfg(vocab, target=False, typ="text"):
    return {"type": typ, "target": target, "vocab_size": vocab, "pad_token_id": 0, "preprocessor_arguments": {}}


def data_config(c):
    return {
        "Formula": _tokcfg(64),
        "IR": {"type": "1D_patches", "target": False, "preprocessor_arguments": {"patch_size": c["ps"]}},
        "Smiles": _tokcfg(c["V"], True),
    }


def model_kwargs(c, dropout=0.1, **over):
    mk = dict(model_type="CustomModel", model_name="facebook/bart-base", d_model=c["d"], num_heads=c["heads"],
              encoder_attention_heads=c["heads"], decoder_attention_heads=c["heads"], encoder_layers=c["layers"],
              decoder_layers=c["layers"], encoder_ffn_dim=c["ffn"], decoder_ffn_dim=c["ffn"], multimodal_norm=True,
              positional_encoding_type="sin_cos", gated_linear=False, max_position_embeddings=1024,
              optimiser="adamw", lr=1e-3, weight_decay=0.0, adam_beta1=0.9, adam_beta2=0.999, dropout=dropout,
              n_beams=10)
    mk.update(over)
    return mk


PAPER = dict(positional_encoding_type="learned", gated_linear=True)  # every IR / mixture / multimodal recipe of the paper


def train_case(name):
    """BASELINE.json configs beyond the headline (SURVEY §8d shapes, no padding): -> dict(data_config, model kwargs,
    batch(B, seed), B, S, T, V, gated, flops_per_sample, workload)."""
    c = dict(C2)
    d, f, L = c["d"], c["ffn"], c["layers"]
    if name == "c2_paper":
        S, T, V, B = 36, 64, 200, 256
        return dict(dc=data_config(c), mk=model_kwargs(c, **PAPER), B=B, S=S, T=T, V=V, gated=True,
                    batch=lambda B, seed: synth_batch(c, B, seed),
                    flops=flops_train(S, T, V, d, f, L, L, c["P"] * c["ps"] * d, True),
                    workload="C2 paper variant: learned pos-enc + GLU FFN (replicate_table_2.sh:16-34), S=15+21x75, T=64, V=200")
    if name in ("c2_postln", "c2_paper_postln"):  # not a BASELINE config: the LN(x + f(x)) layer order (scripts/case_bench.py)
        S, T, V, B = 36, 64, 200, 256
        gated = name == "c2_paper_postln"
        return dict(dc=data_config(c), mk=model_kwargs(c, post_layer_normalisation=False, **(PAPER if gated else {})),
                    B=B, S=S, T=T, V=V, gated=gated, batch=lambda B, seed: synth_batch(c, B, seed),
                    flops=flops_train(S, T, V, d, f, L, L, c["P"] * c["ps"] * d, gated),
                    workload="C2 shapes with post_layer_normalisation=False" + (" (learned + GLU)" if gated else ""))
    if name == "c3":
        S, T, V, B = 14, 24, 64, 128
        dc = {"Formula": _tokcfg(64), "Phosphor_NMR": {"type": "1D_patches", "target": False, "preprocessor_arguments":
                                                        {"patch_size": 1, "encoding_type": "linear_2_layer"}},
              "Smiles": _tokcfg(V, True)}

        def batch(B, seed):
            g = torch.Generator().manual_seed(seed)
            enc = {"Formula": torch.randint(4, 64, (13, B), generator=g), "Phosphor_NMR": torch.randn(1, B, 1, generator=g)}
            return _wire(enc, S, T, V, B, g)
        return dict(dc=dc, mk=model_kwargs(c), B=B, S=S, T=T, V=V, gated=False, batch=batch,
                    flops=flops_train(S, T, V, d, f, L, L, 1 * (d // 2) + (d // 2) * d, False),
                    workload="C3 31P-NMR (1 value, 2-layer patch MLP) + formula -> SMILES (phosphor_from_scratch.sh:37-49), "
                             "S=13+1, T=24, V=64, batch 128")
    if name == "c4":
        S, T, V, B = 16 + 120 + 40 + 23, 96, 300, 128
        dc = {"Formula": _tokcfg(64), "Multiplets": _tokcfg(2048, typ="multiplets"), "Carbon": _tokcfg(2304, typ="carbon"),
              "IR": {"type": "1D_patches", "target": False, "preprocessor_arguments": {"patch_size": 75}},
              "Smiles": _tokcfg(V, True)}

        def batch(B, seed):
            g = torch.Generator().manual_seed(seed)
            enc = {"Formula": torch.randint(4, 64, (16, B), generator=g),
                   "Multiplets": torch.randint(4, 2048, (120, B), generator=g),
                   "Carbon": torch.randint(4, 2304, (40, B), generator=g), "IR": torch.randn(23, B, 75, generator=g)}
            return _wire(enc, S, T, V, B, g)
        return dict(dc=dc, mk=model_kwargs(c, **PAPER), B=B, S=S, T=T, V=V, gated=True, batch=batch,
                    flops=flops_train(S, T, V, d, f, L, L, 23 * 75 * d, True),
                    workload="C4 multimodal formula + 1H multiplets + 13C + IR -> SMILES (configs/data/multimodal/multimodal.yaml), "
                             "learned + GLU, S=16+120+40+23=199, T=96, V=300, batch 128")
    raise KeyError(name)


def _wire(enc, S, T, V, B, g):
    t = torch.randint(4, V, (T + 1, B), generator=g)
    t[0] = 2
    return {"encoder_input": enc, "encoder_pad_mask": torch.zeros(S, B, dtype=torch.bool),
            "decoder_input": {"Smiles": t[:-1].contiguous()}, "decoder_pad_mask": torch.zeros(T, B, dtype=torch.bool),
            "target": t[1:].contiguous()}


class Tok:
    def __init__(self, v):
        self.vocab_size, self.pad_token_id, self.bos_token_id, self.eos_token_id = v, 0, 2, 3

    def batch_decode(self, seqs, skip_special_tokens=True):
        return [" ".join(str(t) for t in s if t > 3) for s in seqs.tolist()]


def synth_batch(c, B, seed, pin=False):
    """Collator wire format (seq-first, True = pad; data/datamodules.py:201-218), no padding, on the host."""
    g = torch.Generator().manual_seed(seed)
    f = torch.randint(4, 64, (c["S_formula"], B), generator=g)
    ir = torch.randn(c["P"], B, c["ps"], generator=g)
    t = torch.randint(4, c["V"], (c["T"] + 1, B), generator=g)
    t[0] = 2
    batch = {
        "encoder_input": {"Formula": f, "IR": ir},
        "encoder_pad_mask": torch.zeros(c["S_formula"] + c["P"], B, dtype=torch.bool),
        "decoder_input": {"Smiles": t[:-1].contiguous()},
        "decoder_pad_mask": torch.zeros(c["T"], B, dtype=torch.bool),
        "target": t[1:].contiguous(),
    }
    if pin:
        batch = map_batch(batch, lambda x: x.pin_memory())
    return batch


def map_batch(b, fn):
    if isinstance(b, dict):
        return {k: map_batch(v, fn) for k, v in b.items()}
    return fn(b) if isinstance(b, torch.Tensor) else b


def batch_bytes(b):
    if isinstance(b, dict):
        return sum(batch_bytes(v) for v in b.values())
    return b.numel() * b.element_size() if isinstance(b, torch.Tensor) else 0


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if sm:
            busy = sorted(sm)[len(sm) // 2:]
            out = {"sm_mhz": sorted(busy)[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# --------------------------------------------------------------------------------------------- CPU (oracle)
def cpu_train_baseline(c, B, steps, warmup):
    """The reference algorithm (oracle port, torch-CPU fp32, dropout 0.1, AdamW + OneCycleLR + clip 1.0) on all host
    cores: spectra/s over `steps` timed steps of batch B."""
    from oracle import spectra_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = orc.OracleConfig(d_model=c["d"], encoder_layers=c["layers"], decoder_layers=c["layers"],
                           encoder_attention_heads=c["heads"], decoder_attention_heads=c["heads"],
                           data_config=data_config(c), dropout=0.1, training=True)
    sd = orc.init_state_dict(cfg, vocab=c["V"], enc_ffn=c["ffn"], dec_ffn=c["ffn"], seed=SEED)
    leaves = []
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith("pos_enc") and ".decoder.embedding." not in k:
            v.requires_grad_(True)
            leaves.append(v)
    opt = torch.optim.AdamW(leaves, lr=1e-3, weight_decay=0.0)
    sch = torch.optim.lr_scheduler.OneCycleLR(opt, 1e-3, total_steps=max(steps + warmup, 10))
    times = []
    for i in range(warmup + steps):
        batch = synth_batch(c, B, SEED + i)
        t0 = time.perf_counter()
        out = orc.wrapper_forward(sd, cfg, batch)
        opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        torch.nn.utils.clip_grad_norm_(leaves, 1.0)
        opt.step()
        sch.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return B * len(times) / total, total / len(times)


def cpu_decode_baseline(c, B, K):
    from oracle import spectra_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = orc.OracleConfig(d_model=c["d"], encoder_layers=c["layers"], decoder_layers=c["layers"],
                           encoder_attention_heads=c["heads"], decoder_attention_heads=c["heads"],
                           data_config=data_config(c))
    sd = orc.init_state_dict(cfg, vocab=c["V"], enc_ffn=c["ffn"], dec_ffn=c["ffn"], seed=SEED)
    batch = synth_batch(c, B, SEED)
    t0 = time.perf_counter()
    with torch.no_grad():
        orc.generate(sd, cfg, batch, n_beams=K)
    dt = time.perf_counter() - t0
    return B / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = dict(C2)
    B = 64
    val, s_per_step = cpu_train_baseline(c, B, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "train spectra/s", "value": val, "unit": "spectra/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(c, B),
        "cpu_baseline": {"value": val, "unit": "spectra/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} timed train steps (fwd+bwd+clip+AdamW) of batch {B}, C2 shapes, fp32"},
        "e2e": {"value": val, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(c, B):
    return {"workload": "C2 IR->SMILES training step (custom_model.yaml d512/6+6/8h/ffn2048, S=15+21 patches of 75, "
                        "T=64, V=200, AdamW+OneCycle+clip1.0, dropout 0.1)",
            "per_gpu_batch": B, "seq_len_enc": c["S_formula"] + c["P"], "seq_len_dec": c["T"],
            "l2": "per-step activations+grads (>1 GB) exceed the 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------------------------- GPU
def time_dominant_gemm(eng, c, B):
    """The dominant kernel class is the tcgen05 GEMM; time its largest instance (decoder FFN-1 with the fused
    bias + GELU + dropout + pre-activation-copy epilogue: M = B*T, N = ffn, K = d) with CUDA events on the launching
    stream.  Ten launches are queued back to back per event pair so the host launch path is not inside the
    measurement; operands + both outputs (153 MB) exceed the 126 MB L2, so no explicit flush is needed."""
    from multimodalanalytical_b200 import ops
    from multimodalanalytical_b200._lib import EPI_GELU
    M, N, K = B * c["T"], c["ffn"], c["d"]
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = eng.W("hf_model.decoder.layers.0.linear1.weight")
    bias = eng.P("hf_model.decoder.layers.0.linear1.bias")
    o1 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    epi = ops.make_epi(EPI_GELU, o1, out2=o2, bias=bias, p_drop=0.1, seed=1, site=3)
    for _ in range(3):
        ops.gemm(a, w, M, N, K, epi)
    torch.cuda.synchronize()
    ts, inner = [], 10
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.gemm(a, w, M, N, K, epi)
        e0.record()
        for _ in range(inner):
            ops.gemm(a, w, M, N, K, epi)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3 / inner)
    t = sorted(ts)[len(ts) // 2]
    return 2.0 * M * N * K / t / 1e12, t, (M, N, K)


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written by scripts/ncu_summary.py); None if not captured."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[key]["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        return None


def run_ours(args):
    import torch.distributed as dist
    from multimodalanalytical_b200 import ops
    from multimodalanalytical_b200.trainer import FusedTrainer
    from multimodalanalytical_b200.wrapper import HFWrapper

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the accelerated path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = dict(C2)
    B = args.batch or c["B"]
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak_src = "measured"
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    else:
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
        peak_src = "fallback"

    PEAKS.update(peaks)
    mk = model_kwargs(c)
    total_steps = args.steps + args.warmup + 8
    model = HFWrapper(data_config=data_config(c), target_tokenizer=Tok(c["V"]), num_steps=4 * total_steps + 16,
                      precision="bf16", seed=SEED, **mk)
    trainer = FusedTrainer(model, clip_grad=1.0, acc_batches=1)
    nb = 4
    host_batches = [synth_batch(c, B, SEED + 1000 * rank + i, pin=True) for i in range(nb)]
    dev_batches = [map_batch(b, lambda x: x.cuda()) for b in host_batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, steps, warmup, e2e):
        for i in range(warmup):
            loss = trainer.train_step(batches[i % nb], i)
            if e2e:
                float(loss)
        barrier()
        l0 = ops.LAUNCHES
        sampler = ClockSampler(local) if (rank == 0 and not e2e) else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss = trainer.train_step(batches[i % nb], i)
            if e2e:
                float(loss)  # device -> host read of the step's result
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) * 1e-3, ops.LAUNCHES - l0, clocks, float(loss)

    t_dev, launches, clocks, last_loss = timed(dev_batches, args.steps, args.warmup, e2e=False)
    t_e2e, _, _, _ = timed(host_batches, args.steps, max(3, args.warmup // 2), e2e=True)
    value = world * B * args.steps / t_dev
    e2e_value = world * B * args.steps / t_e2e
    fl = flops_per_sample_train(c)
    step_tflops = value / world * fl / 1e12

    line = {
        "metric": "train spectra/s", "value": value, "unit": "spectra/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(c, B),
        "e2e": {"value": e2e_value, "unit": "spectra/s", "h2d_bytes_per_step": batch_bytes(host_batches[0]),
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "last_loss": last_loss,
    }
    if world > 1:
        if trainer.peer is not None:
            nvls = any(trainer.peer.mc.values())
            line["gradient_exchange"] = ("peer-memory reduce-scatter + rank-sharded Adam + bf16 all-gather, own kernels over "
                                         + ("NVLink/NVSwitch multicast (multimem.ld_reduce / multimem.st)" if nvls
                                            else "NVLink P2P loads / stores"))
        else:
            line["gradient_exchange"] = "bucketed NCCL all-reduce overlapped with backward"
    # a >= 2.5 s leg: the 20-step headline lasts ~0.15 s, i.e. it runs in the GPU's burst regime
    n_long = max(args.steps, int(2.5 / (t_dev / args.steps)) + 1)
    t_long, _, clocks_long, _ = timed(dev_batches, n_long, 0, e2e=False)
    line["sustained"] = {"steps": n_long, "seconds": t_long, "value": world * B * n_long / t_long, "unit": "spectra/s",
                         "clocks": clocks_long}
    dec = sweep = cfgs = None
    dd = dist if world > 1 else None
    if not args.no_decode:
        # inference shards independent spectra over the ranks with no collective: every rank decodes its
        # own --decode-batch spectra; the aggregate is all molecules / the slowest rank's time
        dec = bench_decode(model, c, args, world, rank, dd)
    model.engine.release_buffers()
    if not args.no_configs:
        cfgs = bench_train_configs(["c2_paper", "c3", "c4"], min(args.steps, 10), world, rank, dd)
    if not args.no_decode and not args.no_sweep:
        sb, sk = [int(x) for x in args.sweep_batches.split(",")], [int(x) for x in args.sweep_beams.split(",")]
        if world > 1 and args.sweep_batches == SWEEP_B and args.sweep_beams == SWEEP_K:
            sb, sk = [1, 256], [10]  # the full curve is a single-GPU measurement; N GPUs decode N independent shards
        sweep = bench_decode_sweep(world, rank, dd, sb, sk)
    if rank == 0:
        line["clocks"] = clocks
        tf, t_k, shape = time_dominant_gemm(model.engine, c, B)
        line["roofline"] = {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": tf / peaks["bf16_tflops"], "traffic": ncu_traffic("gemm2_ffn1_gelu"),
                            "kernel": "tc2::gemm2_kernel<K-major B, EPI_GELU, bf16 out> (cta_group::2, 256x256 tiles)",
                            "shape_MNK": shape, "us_per_launch": t_k * 1e6, "peak_source": f"{peak_src} burst"}
        line["step_roofline"] = {"bound": "tensor", "achieved": step_tflops, "peak": peaks["bf16_tflops_sustained"],
                                 "unit": "TFLOP/s", "frac": step_tflops / peaks["bf16_tflops_sustained"],
                                 "flops_per_sample": fl, "peak_source": f"{peak_src} sustained"}
        line["step_roofline"]["note"] = ("the timed region lasts %.2f s (burst regime); `sustained` repeats it for >= 2.5 s"
                                         % t_dev)
        if world == 1:
            line["pipeline"] = bench_pipeline(trainer, c, B, args.steps)
        if cfgs is not None:
            line["configs"] = cfgs
        if dec is not None:
            line["decode"] = dec
            if sweep is not None:
                dec["sweep_c5"] = sweep
            if world == 1:
                line["guided_decode"] = bench_guided(model, c)
        if world == 1 and not args.no_cpu:
            cb = 64
            cv, cs = cpu_train_baseline(c, cb, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": cv, "unit": "spectra/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"3 timed train steps of batch {cb} (C2 shapes, fp32, oracle port of the "
                                              "reference algorithm; the reference tree does not travel to the GPU box)"}
            if dec is not None:
                # the reference's decode: transformers generate(num_beams=10, use_cache=False) = the oracle's cache-less
                # beam search (wrapper.py:443-451), all 127 steps, on a bounded sample of 2 spectra
                dv, ds = cpu_decode_baseline(c, 2, 10)
                dec["cpu_baseline"] = {"value": dv, "unit": "molecules/s", "cores": torch.get_num_threads(), "kind": "port",
                                       "sample": f"beam-10, 2 spectra, 127 steps without KV cache, {ds:.1f} s (C2 shapes, fp32)"}
        emit(line)
    if world > 1:
        # the captured step graphs hold NCCL work; tearing the communicator down underneath them can block, so leave
        # without the teardown once every rank is done
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def bench_pipeline(trainer, c, B, steps):
    """the batch is assembled on the device from an HBM-resident pre-tokenised dataset; the host only
    sends B sample indices.  Synthetic set in the C2 raw format (1791-point spectra, interpolation + 21 patches of 75
    on the device; 15 formula tokens; 65 target tokens).  Reports the collate time alone and the training throughput
    when every step's batch comes from `DeviceDataset.collate` (H2D per step = B int32 indices)."""
    import numpy as np
    from multimodalanalytical_b200.pipeline import Column, DeviceDataset, HostDataset, IndexSampler, Ragged

    n = 8192
    rng = np.random.default_rng(SEED)
    raw = rng.standard_normal((n, 1791), dtype=np.float32)
    ftok = rng.integers(4, 64, size=(n, c["S_formula"]), dtype=np.int32)
    ttok = rng.integers(4, c["V"], size=(n, c["T"] + 1), dtype=np.int32)
    ttok[:, 0] = 2
    cols = {
        "Formula": Column("tokens", pad_len=c["S_formula"], max_len=c["S_formula"], pad_id=0,
                          tokens=Ragged.from_rows(list(ftok), np.int32)),
        "IR": Column("patches", raw=raw, missing=np.zeros(n, dtype=np.uint8),
                     patch=dict(patch_size=c["ps"], mean=0.0, std=1.0, interpolation=True, overlap=1, masking=False)),
    }
    target = Column("tokens", pad_len=None, max_len=c["T"] + 1, pad_id=0, tokens=Ragged.from_rows(list(ttok), np.int32))
    ds = DeviceDataset(HostDataset(cols, "Smiles", target, n))
    batches = list(IndexSampler(n, B, shuffle=True, seed=SEED, drop_last=True))
    for i in range(3):
        ds.collate(batches[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = min(len(batches), 32)
    e0.record()
    for i in range(reps):
        ds.collate(batches[i])
    e1.record()
    torch.cuda.synchronize()
    t_col = e0.elapsed_time(e1) * 1e-3 / reps
    for i in range(4):
        trainer.train_step(ds.collate(batches[i]), i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        loss = trainer.train_step(ds.collate(batches[i % len(batches)]), i)
    float(loss)
    e1.record()
    torch.cuda.synchronize()
    t_step = e0.elapsed_time(e1) * 1e-3 / steps
    out_bytes = B * (c["P"] * c["ps"] * 4 + c["P"] + 2 * c["S_formula"] * 8 + c["S_formula"] + c["T"] * (8 + 8 + 1))
    in_bytes = B * (c["P"] * c["ps"] * 4 + (c["S_formula"] + c["T"] + 1) * 4)
    return {"dataset_samples": n, "resident_mb": ds.bytes_resident() / 1e6, "collate_us_per_batch": t_col * 1e6,
            "collate_spectra_per_s": B / t_col, "collate_algorithmic_gbs": (in_bytes + out_bytes) / t_col / 1e9,
            "train_spectra_per_s_index_fed": B / t_step, "h2d_bytes_per_step": B * 4}


def bench_guided(model, c, B=64, K=10):
    """formula-guided beam-10 decode (guide fused into the step kernel, host chemistry overlapped with the
    decoder forward).  Synthetic SMILES-like vocabulary; the chemistry backend is a plain element counter, so the number
    shows the cost of the guided decode LOOP (per-step D2H of parent / token ids, host string bookkeeping + memo, one
    chemistry call per NEW hypothesis, H2D of the counts), not of rdkit.  With random weights almost every hypothesis
    of every step is new, so this is the memo's worst case: rows x chemistry calls per step on the host.  The search
    also ends early here - the guide bans <eos> until the formula matches and beams that overshoot die."""
    import re

    from multimodalanalytical_b200.guided import ChemBackend, GuidedFormulaProcessor

    pieces = ["C", "c", "N", "O", "(", ")", "=", "1", "Cl", "Br", "S", "n", "F", "#", "2", "o", "s", "P", "I"]
    vocab = {"<pad>": 0, "<unk>": 1, "<bos>": 2, "<eos>": 3}
    for i in range(4, c["V"]):
        vocab[pieces[(i - 4) % len(pieces)] + ("" if i - 4 < len(pieces) else f"@{i}")] = i  # unique keys, same elements
    elem = re.compile(r"Cl|Br|[CNOSPFI]|[cnosp]")

    class Counter(ChemBackend):
        calls, seconds = 0, 0.0

        def canonical(self, smiles):
            return smiles

        def formula(self, smiles):
            t0 = time.perf_counter()
            n = {}
            for t in elem.findall(smiles):
                t = t if t[0].isupper() else t.upper()
                n[t] = n.get(t, 0) + 1
            out = "".join(f"{k}{v}" for k, v in n.items())
            Counter.calls += 1
            Counter.seconds += time.perf_counter() - t0
            return out

    class VTok:
        vocab_size, pad_token_id, bos_token_id, eos_token_id = c["V"], 0, 2, 3

        def __init__(self):
            self.vocab = vocab

    model.eval()
    batch = map_batch(synth_batch(c, B, SEED + 11), lambda x: x.cuda())
    formulas = ["C30N8O8S4P2F4Cl4Br4I2"] * B

    def run(guided):
        procs = [GuidedFormulaProcessor(K, formulas, VTok(), chem=Counter())] if guided else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = model.generate(batch, n_beams=K, logits_processor=procs)
        e1.record()
        torch.cuda.synchronize()
        steps = model.generator.last_steps if guided else int(out.shape[1]) - 1
        return e0.elapsed_time(e1) * 1e-3, steps

    run(True)
    run(False)
    Counter.calls, Counter.seconds = 0, 0.0
    tg, sg = run(True)
    tu, su = run(False)
    return {"batch": B, "beams": K, "guided_molecules_per_s": B / tg, "guided_steps": sg, "guided_ms_per_step": tg * 1e3 / max(sg, 1),
            "chemistry_calls_per_step": Counter.calls / max(sg, 1), "chemistry_ms_per_step": Counter.seconds * 1e3 / max(sg, 1),
            "unguided_molecules_per_s": B / tu, "unguided_steps": su, "unguided_ms_per_step": tu * 1e3 / max(su, 1),
            "chemistry": "element counter (no rdkit in the image)"}


def decode_floor_s(B, K, steps, S, V, gated, d=512, f=2048, Ld=6):
    """Cached-decode lower bound: per step max(bytes / HBM, flops / tensor) with bytes = decoder weights once
    + self-attention K/V history per row + new K/V per row + cross K/V per spectrum + logits."""
    R = B * K
    g = 1 if gated else 0
    w = Ld * (6 * d * d + (2 + g) * d * f) * 2 + d * V * 2
    tot = 0.0
    for t in range(1, steps + 1):
        by = w + R * Ld * 2 * t * d * 2 + R * Ld * 2 * d * 2 + B * Ld * 2 * S * d * 2 + R * V * 4
        fl = 2.0 * R * (Ld * (6 * d * d + (2 + g) * d * f + 2 * (t + S) * d) + d * V)
        tot += max(by / (PEAKS["hbm_gbs"] * 1e9), fl / (PEAKS["bf16_tflops_sustained"] * 1e12))
    return tot


def time_generate(model, batch, K, reps, dist=None):
    from multimodalanalytical_b200 import ops
    out = model.generate(batch, n_beams=K)  # warm-up + graph capture
    torch.cuda.synchronize()
    ts = []
    l0 = ops.LAUNCHES
    for _ in range(reps):  # each call timed on its own; the median is reported
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = model.generate(batch, n_beams=K)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = sorted(ts)[len(ts) // 2]
    if dist is not None:  # slowest rank
        tt = torch.tensor([t], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt)
    return t, int(out.shape[1]) - 1


def bench_decode(model, c, args, world=1, rank=0, dist=None):
    """Secondary metric: beam-10 molecules/s (KV-cached, CUDA-graph replayed step); random-init weights never emit
    EOS early, so every hypothesis runs the full 127 steps.  Headline at --decode-batch spectra per GPU on the C2 model;
    `roofline_frac` compares the measured time with the cached-decode byte / FLOP floor."""
    K = 10
    model.eval()
    S, V = c["S_formula"] + c["P"], c["V"]
    batch = map_batch(synth_batch(c, args.decode_batch, SEED + 7 + 100 * rank), lambda x: x.cuda())
    t, steps = time_generate(model, batch, K, 3, dist)
    return {"metric": "beam-10 decode molecules/s", "value": world * args.decode_batch / t, "unit": "molecules/s",
            "n_gpus": world, "batch_per_gpu": args.decode_batch, "beams": K, "steps": steps, "ms_per_batch": t * 1e3,
            "ms_per_step": t * 1e3 / steps, "dtype": "bf16", "sharding": "independent spectra per rank, no collective",
            "model": "C2 (custom_model.yaml, sin_cos, no gate)",
            "roofline_frac": decode_floor_s(args.decode_batch, K, steps, S, V, False) / t}


def bench_decode_sweep(world, rank, dist, batches, beams):
    """C5: IR-of-mixtures model (replicate_table_4.sh: learned pos-enc + GLU, IR [B,24,75] + formula [B,16],
    S = 40; custom_model_align weights, head unused while generating), beam 10 and 30, batch sweep per GPU; spectra are
    sharded over the ranks with no collective, so N GPUs decode N x B spectra in the time of the slowest rank."""
    from multimodalanalytical_b200.wrapper import HFWrapper
    c = dict(C2, P=24, S_formula=16, V=120)
    ac = dict(align_network="mlp", hidden_dimension=256, conv_channels=512, kernel_size=5, output_dimension=1800,
              loss_lambda=5, loss_function="mse")
    model = HFWrapper(data_config=data_config(c), target_tokenizer=Tok(c["V"]), num_steps=100, precision="bf16", seed=SEED,
                      **model_kwargs(c, align_config=ac, **PAPER))
    model.eval()
    S, V = c["S_formula"] + c["P"], c["V"]
    rows = []
    for K in beams:
        for B in batches:
            batch = map_batch(synth_batch(c, B, SEED + 13 + 100 * rank), lambda x: x.cuda())
            t, steps = time_generate(model, batch, K, 2, dist)
            st = model.generator._states.get((B, K, model.generation_config["max_length"]))
            plan = model.generator._persist_plan(st) if st is not None else None
            rows.append({"batch_per_gpu": B, "beams": K, "molecules_per_s": world * B / t, "ms_per_step": t * 1e3 / steps,
                         "steps": steps, "roofline_frac": decode_floor_s(B, K, steps, S, V, True) / t,
                         "step": "one launch (%d clusters x 16 CTAs)" % plan[1] if plan else "per-op launches"})
            model.engine.release_buffers()  # K/V caches of B x K rows are shape-keyed workspaces: 48 GB at 1024 x 30
    del model
    torch.cuda.empty_cache()
    return {"workload": "C5 IR mixtures decode: learned + GLU, S=16+24x75, V=120, max_length 128, bf16", "n_gpus": world,
            "rows": rows}


def bench_train_configs(names, steps, world, rank, dist):
    """Per-config training throughput (VERDICT r1 row d+): spectra/s, ms/step and the whole-step tensor-roofline fraction
    for the paper's C2 variant, C3 and C4, measured like the headline (device-resident batches, CUDA events, max over ranks)."""
    from multimodalanalytical_b200.trainer import FusedTrainer
    from multimodalanalytical_b200.wrapper import HFWrapper
    out = {}
    for name in names:
        tc = train_case(name)
        model = HFWrapper(data_config=tc["dc"], target_tokenizer=Tok(tc["V"]), num_steps=1000, precision="bf16",
                          seed=SEED, **tc["mk"])
        tr = FusedTrainer(model, clip_grad=1.0)
        B = tc["B"]
        bt = [map_batch(tc["batch"](B, SEED + 1000 * rank + i), lambda x: x.cuda()) for i in range(2)]
        for i in range(5):
            tr.train_step(bt[i % 2], i)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss = tr.train_step(bt[i % 2], i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        t = float(ms) * 1e-3
        v = world * B * steps / t
        tf = v / world * tc["flops"] / 1e12
        out[name] = {"workload": tc["workload"], "value": v, "unit": "spectra/s", "per_gpu_batch": B,
                     "ms_per_step": t / steps * 1e3, "steps": steps, "flops_per_sample": tc["flops"],
                     "step_roofline_frac": tf / PEAKS["bf16_tflops_sustained"], "last_loss": float(loss)}
        model.engine.release_buffers()
        del tr, model
        torch.cuda.empty_cache()
    return out


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly one JSON line: keep a private duplicate of fd 1 for it and point fd 1 at stderr, so that
    nothing a library prints (NCCL's version banner goes to stdout at every NCCL_DEBUG level >= VERSION) can land there."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--decode-batch", type=int, default=256)
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2-paper / C3 / C4 training legs")
    ap.add_argument("--no-sweep", action="store_true", help="skip the C5 decode batch x beams sweep")
    ap.add_argument("--sweep-batches", default=SWEEP_B)
    ap.add_argument("--sweep-beams", default=SWEEP_K)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
