#!/usr/bin/env python
"""bench.py — MedViLL MLM+ITM pre-training throughput on N B200s (one process per GPU), plus the roofline of the
dominant kernel family, the end-to-end number through the public API with host buffers, and the CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 64] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" = one full pre-training step of BASELINE.json configs[1]: ResNet-50 trunk (cuDNN, frozen, train-mode BN) ->
region gather + projection -> joint embedding -> 12 BERT layers (tcgen05 GEMMs + fused masked attention, BAR mask,
dropout 0.1 on) -> pooler/ITM + MLM heads + both CE losses -> full backward -> (bucketed NCCL all-reduce) -> AdamW.
Synthetic data (seeded), random-init weights, bf16 compute with fp32 master weights.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_JSON_OUT = sys.stdout
METRIC = "MLM+ITM pretrain samples/sec"
UNIT = "samples/s"
# dense algorithmic FLOPs per sample of one training step: encoder linears + attention, fwd + dgrad + wgrad
ENC_FLOPS_PER_SAMPLE = 2.4321e11


def workload(args, n):
    return {"workload": "MedViLL pretrain step (MLM+ITM), BERT-base, Bidirectional Auto-Regressive mask, batch %d/GPU, synthetic "
                        "512x512 CXR + random report tokens, N=180 regions, S=253 (L=436), dropout 0.1, AdamW" % args.batch,
            "global_batch": args.batch * n, "joint_len": 436, "parallelism": "dp%d" % n,
            "images": "uint8 [B,3,512,512] on the wire, ToTensor+Normalize fused on the device (mv_normalize_u8)",
            "l2": "no flush needed: one step streams ~10 GB of saved activations (>> 126 MB L2)"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.stop_flag = [], set(), None, False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown"}
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        import statistics

        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


REF_SAMPLE = ("B = 2 of the step's samples per CPU step (BASELINE.json configs[0]: BERT-base, L = 436, Bidirectional Auto-Regressive mask, "
              "512x512 synthetic CXR, random-init weights, fp32, model.train() with dropout 0.1, forward + CE losses + backward + "
              "HF-3.x AdamW), all host threads; samples/s = 2 / median step time")


def _reference_modules_step_rate(steps, warmup, batch):
    """The reference's OWN modules (models/cxrbert_origin.py + models/image.py, vendored byte for byte into oracle/_ref by
    oracle/make_ref.py) under oracle/ref_shim.py: CXRBERT.forward -> CE(ignore_index=-100) + CE -> backward -> HF-3.x AdamW
    restatement (models/train_origin.py:106-131), fp32, model.train(), dropout on."""
    import types

    import numpy as np
    import torch
    import torch.nn as nn

    os.environ["MEDVILL_REFERENCE"] = os.path.join(ROOT, "oracle", "_ref")
    from oracle import medvill_oracle as orc
    from oracle import ref_shim

    ref_shim.REF_ROOT = os.environ["MEDVILL_REFERENCE"]
    # the reference hard-codes .cuda() on a few index tensors (cxrbert_origin.py:92-95,115,117; image.py:60); this arm is
    # the CPU path, so on a box WITH a GPU those calls are made no-ops for the duration of the timing and restored after
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        return _reference_modules_timed(steps, warmup, batch, ref_shim, orc, types, np, torch, nn)
    finally:
        torch.Tensor.cuda = orig_cuda


def _reference_modules_timed(steps, warmup, batch, ref_shim, orc, types, np, torch, nn):
    cxr, _ = ref_shim.load_reference_models()
    from transformers import BertConfig

    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    cfg = orc.Cfg()
    ref_shim.set_bert_config()
    margs = types.SimpleNamespace(bert_model="bert-base-scratch", img_hidden_sz=2048, embedding_size=768, hidden_size=768,
                                  dropout_prob=0.1, img_postion=True, img_encoder="random-pixel", num_image_embeds=cfg.num_image_embeds,
                                  img_size=cfg.img_size, disturbing_mask=False, vocab_size=cfg.vocab)
    model = cxr.CXRBERT(BertConfig.from_pretrained("bert-base-uncased"), margs).train()
    params = [p for p in model.parameters() if p.requires_grad]
    state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in params]
    mlm_crit, itm_crit = nn.CrossEntropyLoss(ignore_index=-100), nn.CrossEntropyLoss()
    times = []
    for i in range(warmup + steps):
        b = orc.synthetic_batch(cfg, batch, seed=1000 + i, mode=orc.MODE_BAR)
        t = lambda k: torch.as_tensor(b[k])
        t0 = time.perf_counter()
        mlm, itm = model(t("cls_tok"), t("input_ids"), t("attn_masks"), t("segment"), b["image"], t("sep_tok"))
        loss = itm_crit(itm, t("is_aligned")) + mlm_crit(mlm.transpose(1, 2), t("txt_labels"))
        for p in params:
            p.grad = None
        loss.backward()
        step = i + 1
        with torch.no_grad():                       # transformers-3.x AdamW, correct_bias=True, weight_decay 0, lr 1e-5
            ss = 1e-5 * (1 - 0.999 ** step) ** 0.5 / (1 - 0.9 ** step)
            for p, (m, v) in zip(params, state):
                if p.grad is None:
                    continue
                m.mul_(0.9).add_(p.grad, alpha=0.1)
                v.mul_(0.999).addcmul_(p.grad, p.grad, value=0.001)
                p.addcdiv_(m, v.sqrt().add_(1e-6), value=-ss)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    assert np.isfinite(float(loss.detach()))
    return batch / med, med, os.cpu_count() or 1


def cpu_reference_step_rate(steps, warmup, batch=2):
    """CPU arm: the reference's own modules when oracle/_ref holds them (kind "reference"), else the oracle port
    (oracle/medvill_oracle.py, pinned against the real reference; kind "port").  Returns (samples/s, median s, cores, kind)."""
    if os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "models", "cxrbert_origin.py")):
        try:
            return _reference_modules_step_rate(steps, warmup, batch) + ("reference",)
        except Exception as e:                      # e.g. a transformers version the shims do not cover: fall back to the port
            sys.stderr.write("bench: reference modules under oracle/_ref failed (%s: %s); timing the oracle port instead\n" % (type(e).__name__, e))
    import torch

    from oracle import medvill_oracle as orc

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = orc.Cfg()
    params = orc.synth_params(cfg, seed=0)
    state = {}
    times = []
    for i in range(warmup + steps):
        batch_i = orc.synthetic_batch(cfg, batch, seed=1000 + i, mode=orc.MODE_BAR)
        t0 = time.perf_counter()
        out = orc.loss_and_grads(params, batch_i, cfg)
        upd = orc.adamw_step({n: params[n] for n in out["grads"]}, out["grads"], state, lr=1e-5, step=i + 1)
        params.update(upd)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return batch / med, med, os.cpu_count() or 1, "port"


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on this box's host cores.  Same metric / unit /
    workload as our arm; each CPU step is a BOUNDED SAMPLE of that workload (2 of its samples) — stated in config.sample and
    cpu_baseline.sample.  Under torchrun only rank 0 runs; the GPU count does not enter (n_gpus is echoed for the driver)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    rate, med, cores, kind = cpu_reference_step_rate(steps, warm)
    n = args.gpus
    cfg = workload(args, n)
    cfg["sample"] = REF_SAMPLE
    cfg["runs_on"] = "host CPU, %d threads; no GPU work in this arm" % cores
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": n, "steps": steps, "warmup": warm,
            "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d timed steps after %d warm-up, median; %s" % (steps, warm, REF_SAMPLE)},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def run_ours(args):
    import types

    import torch

    import medvill_b200  # noqa: F401
    from medvill_b200 import _lib
    from medvill_b200.config import BertConfig
    from medvill_b200.data.prefetch import DevicePrefetcher
    from medvill_b200.data.synthetic import as_tuple, synthetic_batch
    from medvill_b200.models import CXRBERT

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libmedvill_sm100 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    B = args.batch
    margs = types.SimpleNamespace(img_hidden_sz=2048, embedding_size=768, hidden_size=768, dropout_prob=args.dropout,
                                  img_encoder="random-pixel", allow_random_trunk=True, num_image_embeds=180, img_size=512, seq_len=253, lr=1e-5,
                                  precision="bf16", max_micro_batch=B, seed=123)
    torch.manual_seed(0)
    model = CXRBERT(BertConfig.from_pretrained("bert-base-uncased"), margs).to(dev)
    model.train()
    eng = model.engine(B)
    if world > 1:
        def bcast(raw):
            box = [raw]
            torch.distributed.broadcast_object_list(box, src=0)
            return box[0]
        eng.comm_init(rank, world, bcast)
        torch.distributed.broadcast(eng.params, src=0)
        model.sync_params()

    # a small pool of distinct synthetic batches (pinned host memory), cycled
    pool = [synthetic_batch(B, seed=123 + 17 * rank + i, pin=True, image_dtype=torch.uint8) for i in range(2)]
    dev_pool = [{k: (v if k == "txt_labels" else v.to(dev)) for k, v in b.items()} for b in pool]

    def step_device(i):
        b = dev_pool[i % len(dev_pool)]
        return model.pretrain_step(b["cls_tok"], b["input_ids"], b["txt_labels"], None, b["image"], b["segment"], b["is_aligned"],
                                   b["sep_tok"], mode=b["mode"], t_len=b["t_len"], lazy=True)

    def run_steps(first, n, step_fn):
        """n steps; every step's loss / accuracy counters are copied to pinned host memory inside the step and consumed
        here one step later (the trainer's logging pattern), so the GPU queue never drains on a host read."""
        pending, out = None, None
        for i in range(first, first + n):
            nxt = step_fn(i)
            if pending is not None:
                out = pending()
            pending = nxt
        return pending() if pending is not None else out

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    out = run_steps(0, args.warmup, step_device)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.lib().mv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run_steps(0, args.steps, step_device)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib().mv_launch_count() - launches0
    sampler.stop_flag = True
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * args.steps / (ms / 1e3)

    # ---- end to end through the public API: pinned host tensors -> prefetcher -> pretrain_step -> loss on the host ----
    class Cycle:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __iter__(self):
            for i in range(self.n):
                yield as_tuple(pool[i % len(pool)])

    pf = DevicePrefetcher(Cycle(0), dev, host_indices=(2,))     # one prefetcher: its two device buffer sets persist across runs

    def run_e2e(n):
        pf.loader = Cycle(n)
        it = iter(pf)

        def step_e2e(i):
            cls_tok, input_ids, txt_labels, attn, img, segment, is_aligned, sep_tok, _ = next(it)
            return model.pretrain_step(cls_tok, input_ids, txt_labels, attn, img, segment, is_aligned, sep_tok,
                                       mode=attn[:, 0].to(torch.uint8), t_len=attn[:, 1].to(torch.int32), lazy=True)
        last = run_steps(0, n, step_e2e)
        return pf.bytes_last, last

    run_e2e(max(3, args.warmup))          # own warm-up: staging buffers, cuDNN plans for the staged tensors
    barrier()
    t0 = time.perf_counter()
    e0.record()
    h2d, last = run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 0.0)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_e2e = float(t.item())
    e2e_value = B * world * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel family (tcgen05 GEMM): CUDA events around every launch, 2 extra steps ----
    import ctypes as C

    _lib.check(_lib.lib().mv_profile(eng._h, 1))
    for i in range(2):
        step_device(i)
    NT = 8
    tms, tfl, tcn = (C.c_double * NT)(), (C.c_double * NT)(), (C.c_int32 * NT)()
    _lib.check(_lib.lib().mv_profile_read(eng._h, tms, tfl, tcn, NT))
    _lib.check(_lib.lib().mv_profile(eng._h, 0))
    GEMM_TAGS = {0: "plain fwd/dgrad", 3: "wgrad (fp32 reduce-add)", 4: "GELU-forward epilogue", 5: "GELU-backward epilogue", 6: "residual epilogue (fp32 stream)"}
    # tag 0 of the three-entry view = the whole tcgen05 GEMM family
    pms = [sum(tms[t] for t in GEMM_TAGS), tms[1], tms[2]]
    pfl = [sum(tfl[t] for t in GEMM_TAGS), tfl[1], tfl[2]]
    pcn = [sum(tcn[t] for t in GEMM_TAGS), tcn[1], tcn[2]]
    gemm_classes = {name: {"launches_per_step": tcn[t] // 2, "ms_per_step": tms[t] / 2,
                           "tflops": (tfl[t] / (tms[t] * 1e-3) / 1e12) if tms[t] > 0 else None} for t, name in GEMM_TAGS.items()}
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained"
    gemm_tf = pfl[0] / (pms[0] * 1e-3) / 1e12 if pms[0] > 0 else 0.0
    step_ms = ms / args.steps
    enc_tf = ENC_FLOPS_PER_SAMPLE * B / (step_ms * 1e-3) / 1e12
    # north_star scope "encoder GEMMs + attention": dense encoder FLOPs over the time of ALL tcgen05 GEMM launches (the
    # encoder's, plus the small image-projection / pooler / MLM-head ones, i.e. a slight under-estimate) + attention
    enc_ms = (pms[0] + pms[1] + pms[2]) / 2
    enc_scope_tf = ENC_FLOPS_PER_SAMPLE * B / (enc_ms * 1e-3) / 1e12 if enc_ms > 0 else 0.0
    # DRAM traffic of the same kernel family, per launch like `achieved`: measured offline with ncu (one pass over a
    # profiled step, dram__bytes_read.sum + dram__bytes_write.sum per launch; tools/summarize_launches.py --json)
    traffic, traffic_src = None, None
    try:
        tr_name = next(n for n in ("r02_step_traffic.json", "r01_step_traffic.json") if os.path.exists(os.path.join(ROOT, "profiles", n)))
        with open(os.path.join(ROOT, "profiles", tr_name)) as f:
            tr = json.load(f)["gemm_tc05"]
        traffic = tr["dram_bytes"] / tr["launches"]
        traffic_src = "profiles/%s: %.1f GB over %d launches of one step (ncu)" % (tr_name, tr["dram_bytes"] / 1e9, tr["launches"])
    except Exception:
        pass
    n_gemm = max(1, pcn[0] // 2)
    roofline = {"bound": "tensor", "kernel": "gemm_tc05_kernel (all %d launches of a step)" % (pcn[0] // 2), "achieved": gemm_tf,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tf / peak_tf, "traffic": traffic, "traffic_unit": "bytes/launch (DRAM)",
                "traffic_source": traffic_src, "algorithmic_flops_per_launch": pfl[0] / 2 / n_gemm, "peak_source": peak_src,
                "gemm_classes": gemm_classes, "gemm_ms_per_step": pms[0] / 2, "attn_fwd_ms_per_step": pms[1] / 2, "attn_bwd_ms_per_step": pms[2] / 2,
                "attn_fwd_tflops_dense": pfl[1] / (pms[1] * 1e-3) / 1e12 if pms[1] > 0 else None,
                "attn_bwd_tflops_dense": pfl[2] / (pms[2] * 1e-3) / 1e12 if pms[2] > 0 else None,
                "encoder_gemm_plus_attention_ms": enc_ms, "encoder_gemm_plus_attention_tflops_dense": enc_scope_tf,
                "encoder_gemm_plus_attention_frac_of_peak": enc_scope_tf / peak_tf,
                "encoder_dense_tflops_over_whole_step": enc_tf, "encoder_frac_of_peak_over_whole_step": enc_tf / peak_tf}

    def teardown():
        # identical shutdown order on every rank: engine (its NCCL communicator) first, then torch.distributed; then
        # leave without running finalizers (a rank that exits while its peer is still inside a collective teardown hangs)
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            torch.distributed.barrier()
            model._release_engine()
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
            os._exit(0)

    if rank != 0:
        teardown()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, med, cores, kind = cpu_reference_step_rate(3, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": "3 timed steps after 1 warm-up, median; " + REF_SAMPLE}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "last_loss": out["loss"]}
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU per step")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="fixed GLOBAL batch split evenly over the ranks (BASELINE.json configs[3]: 512 -> 256 / 128 / 64 per GPU at 2 / 4 / 8 "
                         "GPUs, one launch sequence per step); overrides --batch and reports scaling 'strong'")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.strong = False
    if args.global_batch > 0:
        world = int(os.environ.get("WORLD_SIZE", 1))
        if args.global_batch % world:
            raise SystemExit("--global-batch %d is not divisible by %d ranks" % (args.global_batch, world))
        args.batch, args.strong = args.global_batch // world, True
    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library chatter)
    # is routed to stderr; the JSON goes to the saved descriptor.
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    wd = int(os.environ.get("MEDVILL_BENCH_WATCHDOG", "0"))
    if wd > 0:      # debugging aid: dump every thread's Python stack and exit if the run is still alive after `wd` seconds
        import faulthandler

        faulthandler.dump_traceback_later(wd, exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
