#!/usr/bin/env python
"""PMGT pre-training throughput benchmark (driver contract: one JSON line on rank 0).

    python bench.py --gpus N --steps K --warmup W            # ours (B200, libpmgt_b200.so)
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on the host CPU cores

Metric (BASELINE.json): node-contexts/s of the pre-training step
{sample -> gather -> encode -> losses -> backward -> allreduce -> AdamW}; one
node-context = one sampled row of node_ids consumed by the encoder (B targets +
B*P pairs per step; the masked re-encode is extra work, not extra contexts).

Workload at N = 1 (BASELINE.json configs[1]): synthetic TG-shaped item graph
(10,834 nodes / 38,252 edges, 1536-d + 768-d features), default encoder
(H = I = 128, 5 layers, 1 head, L = 6, hops [16, 8, 4], P = 10), bf16 tensor-core
GEMMs, 4096 targets (45,056 contexts) per GPU per step; weak scaling for N > 1
(per-GPU batch fixed, targets sharded by rank, one NCCL allreduce of the flat
gradient per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="TG", choices=["VG", "TG", "1M"])
    ap.add_argument("--encoder", default="default", choices=["default", "wide"],
                    help="wide = BASELINE config 5: H=768, 12 layers, 12 heads, I=3072, 32 neighbours (L=33)")
    ap.add_argument("--batch", type=int, default=4096, help="targets per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel-family event profile to this JSON file")
    return ap.parse_args()


def load_traffic(workload, batch):
    """Measured DRAM bytes per launch of each kernel family (one ncu --set full capture, summarised in profiles/)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        d = json.load(f)
    if d.get("workload") != workload or d.get("targets_per_gpu_per_step") != batch:
        return {}
    return d.get("traffic_bytes_per_launch", {})


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback")  # B200_PROFILING.md fallback


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores, in a subprocess
# ---------------------------------------------------------------------------
def run_cpu_port(workload, steps, warmup, budget_s):
    cmd = [sys.executable, "-m", "oracle.cpu_baseline", "--workload", workload, "--steps", str(steps), "--warmup",
           str(warmup), "--budget-s", str(budget_s)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    if r.returncode != 0:
        raise RuntimeError("cpu baseline failed: " + r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = "TG" if a.workload == "1M" else a.workload  # the CPU port cannot hold the 1M graph's nx-free arrays cheaply
    d = run_cpu_port(wl, a.steps, a.warmup, budget_s=150.0)
    sample = (f"{d['targets_per_step']} targets ({int(d['contexts'] / d['steps'])} contexts) per step x {d['steps']} steps on "
              f"the {wl} graph; sampler = process pool over {d['cores']} cores, model fp32 torch with {d['cores']} threads")
    line = {
        "impl": "reference", "metric": "pmgt_pretrain_node_contexts_per_s", "value": d["value"], "unit": "contexts/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": d["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, d["targets_per_step"], "cpu"),
        "cpu_baseline": {"value": d["value"], "unit": "contexts/s", "cores": d["cores"], "kind": "port", "sample": sample,
                         "sampler_contexts_per_s": d["sampler_contexts_per_s"], "model_contexts_per_s": d["model_contexts_per_s"]},
        "e2e": {"value": d["value"], "unit": "contexts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


ENCODERS = {
    "default": (dict(), "default encoder H=128 I=128 5 layers 1 head, L=6"),
    "wide": (dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, max_ctx_neigh=32),
             "wide encoder H=768 I=3072 12 layers 12 heads, L=33"),
}


def workload_config(wl, batch, where, encoder="default"):
    from pmgt_b200 import synthetic
    n, m, _, _ = synthetic.SHAPES[wl]
    return {"workload": f"PMGT pre-training step on synthetic {wl}-shaped item graph ({n} nodes / {m} edges, 1536-d visual + "
                        f"768-d text features), {ENCODERS[encoder][1]}, hops [16,8,4], 10 pairs/target",
            "targets_per_gpu_per_step": batch, "contexts_per_gpu_per_step": batch * 11, "where": where,
            "l2": "per-step working set (activations, several GB at the default batch) is >> the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nme, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws != a.gpus:
        if ws == 1 and a.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pmgt_b200 has no CPU fallback (use --impl reference for the CPU port)")

    # CPU baseline first (rank 0, N = 1 only), in a CUDA-free subprocess, so it never overlaps the GPU timing
    cpu = None
    if rank == 0 and ws == 1 and not a.no_cpu_baseline:
        wl = "TG" if a.workload == "1M" else a.workload
        d = run_cpu_port(wl, steps=2, warmup=0, budget_s=20.0)
        cpu = {"value": d["value"], "unit": "contexts/s", "cores": d["cores"], "kind": "port",
               "sample": f"{d['targets_per_step']} targets x {d['steps']} steps of the same {wl} workload "
                         f"(sampler pool {d['cores']} procs: {d['sampler_contexts_per_s']:.0f} ctx/s; fp32 torch model "
                         f"{d['cores']} threads: {d['model_contexts_per_s']:.0f} ctx/s)"}

    torch.cuda.set_device(local_rank)
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from pmgt_b200 import ops, synthetic, trainer

    dev = torch.device("cuda", local_rank)
    args = trainer.make_args(synthetic=a.workload, train_batch_size=a.batch, seed=0, **ENCODERS[a.encoder][0])
    args.device = dev
    trainer.set_seed(0)
    args.graph, args.feat_init_emb = trainer._load_graph_and_features(args)
    trainer.init_dataloader(args)
    trainer.init_model(args)
    tm = trainer.PMGTTrainerModel(args)
    ds = args.train_dataset
    B = a.batch
    n_train = len(ds)

    def step_indices(step):
        # weak scaling: every rank takes its own B targets of a (seed, step)-derived permutation (with wrap-around:
        # TG has 8,667 training nodes, so a 4096-target step per rank re-visits nodes across ranks / steps)
        perm = trainer.epoch_permutation(n_train, 0, step)
        reps = (B * ws + n_train - 1) // n_train
        perm = np.concatenate([perm] * reps) if reps > 1 else perm
        return perm[rank * B: rank * B + B].astype(np.int64)

    total = a.warmup + a.steps
    idx_host = [torch.from_numpy(step_indices(s)).pin_memory() for s in range(total + a.steps)]
    idx_dev = [t.to(dev) for t in idx_host[:total]]

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(src, first, n, read_loss=False, prefetch=True):
        """n consecutive steps; the batch of step k+1 (sampling + NFR corruption, and for host-resident indices their
        H2D copy) is prepared on the trainer's side stream while step k runs -- what a DataLoader's prefetch does."""
        loss = None
        for k in range(n):
            s = first + k
            loss = tm.train_on_indices(ds, src[s], epoch=s)
            if prefetch and k + 1 < n:
                tm.prefetch(ds, src[s + 1], epoch=s + 1)
            if read_loss:
                loss = tm.last_loss()  # D2H read of this step's loss (pinned buffer; waits for the forward pass only)
        return loss

    # ---- device-resident run: `value`
    run_steps(idx_dev, 0, a.warmup)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = ops.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(idx_dev, a.warmup, a.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.LAUNCHES[0] - launches0
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    contexts_per_step = B * 11 * ws
    value = contexts_per_step * a.steps / (ms / 1e3)

    # ---- end-to-end run: host index batch (pinned) -> H2D -> step -> loss D2H, every step
    barrier()
    w0 = time.perf_counter()
    e0.record()
    loss_host = run_steps(idx_host, total, a.steps, read_loss=True)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
    t = torch.tensor([ms_e2e], device=dev)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = contexts_per_step * a.steps / (float(t) / 1e3)

    # ---- per-kernel-family profile (CUDA events on the launching stream, inside real steps) -> roofline
    roof, prof = None, None
    # every rank runs the profiled steps (they contain the gradient allreduce); only rank 0 records events
    n_prof = min(3, a.steps)
    if rank == 0:
        ops.PROFILE = []
    # (no side-stream prefetch and no programmatic dependent launch here: every kernel is timed alone)
    pdl_was = ops.set_pdl(False)
    run_steps(idx_dev, a.warmup, n_prof, prefetch=False)
    barrier()
    ops.set_pdl(pdl_was)
    if rank == 0:
        recs, ops.PROFILE = ops.PROFILE, None
        prof = ops.profile_summary(recs)
        tot = sum(d["ms"] for d in prof.values())
        for d in prof.values():
            d["ms_per_step"] = d["ms"] / n_prof
            d["share"] = d["ms"] / tot
            d["gbs"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
            d["tflops"] = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
        top = max(prof, key=lambda k: prof[k]["ms"])
        peaks = load_peaks()
        d = prof[top]
        # the sampler's algorithmic bytes come from the kernel's visited-degree output
        if top == "sample_contexts":
            from pmgt_b200.datasets import context_keys, sample_contexts
            r = torch.as_tensor(ds.node_ids[:4096], device=dev)
            _, _, vdeg = sample_contexts(ds.item_graph, r, context_keys(0, r, 0), ds.hop_sampling_sizes, 5, 0, True)
            per_ctx = float(vdeg.float().mean()) * 8 + 145 * 8 + 6 * 12
            d["gbs"] = per_ctx * contexts_per_step / ws / (d["ms_per_step"] * 1e-3) / 1e9
        roof = {"kernel": top, "bound": "hbm", "achieved": d["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": d["gbs"] / peaks["hbm_gbs"], "traffic": load_traffic(a.workload, B).get(top),
                "algorithmic_bytes_per_launch": d["bytes"] / d["calls"], "peak_source": peaks["source"],
                "share_of_step": d["share"], "launches_per_step": d["calls"] / n_prof,
                "avg_launch_us": 1e3 * d["ms"] / d["calls"], "tensor_tflops": d["tflops"]}
        if a.profile_out:
            with open(a.profile_out, "w") as f:
                json.dump({"per_kernel_family": prof, "steps_profiled": n_prof, "sum_ms_per_step": tot / n_prof,
                           "ms_per_step_timed": ms / a.steps}, f, indent=1)

    if rank == 0:
        line = {
            "metric": "pmgt_pretrain_node_contexts_per_s", "value": value, "unit": "contexts/s", "n_gpus": ws,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(a.workload, B, "gpu", a.encoder),
            "e2e": {"value": e2e_value, "unit": "contexts/s", "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": 4,
                    "api": "pmgt_b200.trainer.PMGTTrainerModel.train_on_indices(pinned host index batch) + .last_loss() every step"},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "loss_last": loss_host,
        }
        print(json.dumps(line), flush=True)
    if ws > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)


if __name__ == "__main__":
    main()
