#!/usr/bin/env python
"""bench.py -- render + warp + photometric forward+backward frames/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One step = one pass of the hot path over one batch of synthetic frame pairs (BASELINE.json configs[2],
which contains configs[1] -- 32 hand+object mesh renders of 9 104 triangles at 256x256 -- as its
rasterizer part): per rank 16 pairs -> 2 x 16 mesh renders (fill_back, 2F = 9104 faces), occlusion
check, 2 x 16 image warps + masked L1, and the full backward (texture AND geometry gradients, i.e.
detach_renders=False so that nothing the reference computes is skipped) down to the camera-space
vertices.  frames/s = 2 * pairs * ranks * steps / time.  Weak scaling: every rank gets its own 16 pairs
and there is no data-path collective.

The JSON line follows the driver contract; `roofline` is for hoc_raster_backward_kernel (the kernel
BASELINE.json's north_star names), timed live with CUDA events around each of its launches in the timed
region; `cpu_baseline` / `--impl reference` time the CPU oracle (the reference has no CPU renderer and
its CUDA extension is not installable here -- DESIGN.md) on the host cores.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PAIRS = 16          # frame pairs per rank and step (configs[2])
SIZE = 256          # raster side of configs[2] (square frames); WIDTH / HEIGHT are the frame size
WIDTH, HEIGHT = 256, 256
CONFIG = 2
N_SETS = 3          # distinct input sets cycled between steps (> L2 between reuse)
METRIC = "render+warp+photometric fwd+bwd frames/sec @256x256"
UNIT = "frames/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        while self.ok and not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.001)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _make_sets(n_sets, pairs, size, device, pin=False, rank=0, world=1):
    """`n_sets` input sets for this rank: the GLOBAL batch of pairs*world frame pairs is generated identically on
    every rank (same seed) and rank r keeps its shard [r*pairs, (r+1)*pairs) -- sharding.shard_range.
    `size`: int (square frames) or (width, height)."""
    import torch
    from handobjectconsist_b200 import sharding, synth

    width, height = (size, size) if isinstance(size, int) else size
    sets = []
    for i in range(n_sets):
        sc = synth.make_scene(pairs * world, width, height, seed=1000 + i)
        lo, hi = sharding.shard_range(pairs * world, rank, world)
        sc = {k: (v[lo:hi].contiguous() if torch.is_tensor(v) else v) for k, v in sc.items()}
        if device is not None:
            sc = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in sc.items()}
        elif pin:
            sc = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in sc.items()}
        sets.append(sc)
    return sets


def _samples_from_scene(sc, hand_v=778):
    """The reference's batch layout: two sample dicts + two result dicts (warpbranch.py:27-44)."""
    from handobjectconsist_b200.queries import BaseQueries, TransQueries

    import torch

    def own(t):  # contiguous (and pinned, for host sets) tensors: slices of pinned storage are not DMA-able as is
        t = t.contiguous()
        return t.pin_memory() if (not t.is_cuda and sc["verts1"].is_pinned()) else t

    obj_faces = own(sc["faces"][:, 1552:] - hand_v)
    samples, results = [], []
    for verts, img, jit in ((sc["verts1"], sc["image_ref"], sc["jitter_mask_ref"]),
                            (sc["verts2"], sc["image"], sc["jitter_mask"])):
        samples.append({TransQueries.IMAGE: img, TransQueries.JITTERMASK: jit, TransQueries.CAMINTR: sc["K"],
                        BaseQueries.OBJFACES: obj_faces, BaseQueries.OBJVERTS3D: own(verts[:, hand_v:]),
                        BaseQueries.HANDVERTS3D: own(verts[:, :hand_v])})
        results.append({"recov_handverts3d": own(verts[:, :hand_v]), "recov_objverts3d": own(verts[:, hand_v:])})
    return samples, results


def _bind_to_gpu_numa_node(local_rank):
    """Pin this process (and so its later pinned-host allocations, first touch) to the CPUs of the NUMA node its GPU
    hangs off: with 8 ranks streaming frames over PCIe, a rank whose staging buffers live on the other socket pays the
    inter-socket hop on every copy.  Best effort -- returns a description or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception:
        return None


def run_native(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from handobjectconsist_b200 import _config, _lib, sharding, warpbranch
    from handobjectconsist_b200.graphed import GraphedConsistStep
    from handobjectconsist_b200.neurender.renderer import Renderer
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: the product path has no CPU fallback")
    numa = _bind_to_gpu_numa_node(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    W, H = WIDTH, HEIGHT
    S = max(W, H)  # the reference renders on the square that contains the frame (warpreg.py:29,40-45)
    renderer = Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                        K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True,
                        near=0.1, no_light=True)
    criterion = PyramidCriterion("l1")
    step_kw = dict(gt_refs=True, first_only=True, use_backward=True, detach_renders=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        """K calls of fn(i) bracketed by barrier + synchronize, device time from CUDA events."""
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            fn(i)
        t1.record()
        barrier()
        return t0.elapsed_time(t1)

    def timed_blocks(fn, steps, min_total_ms, max_blocks=300):
        """The K-step block above, repeated until ~min_total_ms of device time has been timed (a 5 ms region moves by
        per cent on one hiccup).  Every block is max-reduced over the ranks; returns the per-block times."""
        first = sharding.max_over_ranks([timed(fn, steps)], device=dev)[0]
        n = int(min(max_blocks, max(1, -(-min_total_ms // max(first, 1e-3)))))
        blocks = [first] + [timed(fn, steps) for _ in range(n - 1)]
        return sharding.max_over_ranks(blocks, device=dev)

    def median(v):
        v = sorted(v)
        return v[len(v) // 2]

    dsets = _make_sets(N_SETS, PAIRS, (W, H), dev, rank=rank, world=world)
    hand_face = dsets[0]["faces"][0, :1552].clone()
    ignore = dsets[0]["hand_ignore_faces"]
    dbatches = [_samples_from_scene(sc) for sc in dsets]

    def eager_step(samples, results, visuals=False):
        hv = results[0]["recov_handverts3d"].detach().requires_grad_(True)
        ov = results[0]["recov_objverts3d"].detach().requires_grad_(True)
        res = [{"recov_handverts3d": hv, "recov_objverts3d": ov}, results[1]]
        loss, _ = warpbranch.forward(samples, res, hand_face, renderer, (W, H), criterion, hand_ignore_faces=ignore,
                                     return_visuals=visuals, **step_kw)
        loss.backward()
        return loss

    # ---- measured coverage of the synthetic scene (the algorithmic-byte formulas below use it) ----
    from handobjectconsist_b200 import consist
    stats = {}
    consist.STATS = stats
    eager_step(*dbatches[0])
    consist.STATS = None
    coverage = float(stats.get("coverage", 0.0))

    # ---- eager arm: every launch issued from python (CPU-launch-bound), for reference ----
    for i in range(args.warmup):
        eager_step(*dbatches[i % N_SETS])
    eager_ms = timed(lambda i: eager_step(*dbatches[i % N_SETS]), args.steps)
    if args.eager_only:
        if rank == 0:
            _emit({"eager_ms_per_step": eager_ms / args.steps})
        return None

    # ---- graphed arm, inputs resident in HBM: the headline `value` ----
    # one captured step per resident input set: the set IS the graph's static input buffers, so a timed step is
    # exactly one graph replay (no copies); the sets rotate so that consecutive steps touch different data
    launches0 = L.hoc_launch_count(-1)

    def capture(k, **kw):
        return GraphedConsistStep(renderer, criterion, (W, H), hand_face, *dbatches[k], hand_ignore_faces=ignore,
                                  warmup=1, **step_kw, **kw)

    gsteps_dev = [capture(0)]
    launches_per_step = int(L.hoc_launch_count(-1) - launches0) // 2  # one warm-up run + the captured run
    gsteps_dev += [capture(k) for k in range(1, N_SETS)]
    gstep = gsteps_dev[0]

    def graphed_step(i):
        gsteps_dev[i % N_SETS].replay()

    for i in range(args.warmup):
        graphed_step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    blocks = timed_blocks(graphed_step, args.steps, min_total_ms=1000.0)
    clocks = sampler.finish()
    ms = median(blocks)
    launches = launches_per_step * args.steps

    # ---- the same step WITH the visualisation returns of pair_consist (warps / diffs / warp_mask), for transparency:
    # `value` times the training step, which does not produce them ----
    vis_ms = None
    if world == 1:
        gvis = capture(0, return_visuals=True)
        for _ in range(args.warmup):
            gvis.replay()
        vis_ms = median(timed_blocks(lambda i: gvis.replay(), args.steps, min_total_ms=200.0))
        del gvis

    # ---- per-kernel device times, measured where the kernels run in production: inside the graph.  A separate,
    # instrumented capture brackets every launch of this library with external event nodes (hoc_timer_*); it is
    # replayed after the timed region so that the event nodes do not perturb `value`.
    timed_mask = sum(1 << v for k, v in _lib.KERNEL_IDS.items() if k not in ("grad_extent", "raster_bwd_group"))
    probe = capture(0, before_capture=lambda: L.hoc_timer_begin(timed_mask))  # arm right before the capture
    L.hoc_timer_pause()
    buf = (ctypes.c_float * 8192)()
    ids = (ctypes.c_int * 8192)()
    id2name = {v: k for k, v in _lib.KERNEL_IDS.items()}
    per_kernel = {}
    probe_steps = max(min(args.steps, 50), 5)
    for i in range(probe_steps):
        probe.load(*dbatches[i % N_SETS])
        probe.replay()
        torch.cuda.synchronize()
        n_k = L.hoc_timer_peek(buf, ids, 8192)
        for j in range(n_k):
            if buf[j] >= 0:
                per_kernel.setdefault(id2name[ids[j]], []).append(buf[j])
    L.hoc_timer_begin(0)
    # the rasterizer backward (the north star's kernel) as ONE bracket: the external event nodes cost ~4 us per pair, so
    # three kernels timed one by one carry ~12 us of instrumentation, the group ~4
    probe_g = capture(0, before_capture=lambda: L.hoc_timer_begin(1 << _lib.KERNEL_IDS["raster_bwd_group"]))
    L.hoc_timer_pause()
    group_ms = []
    for i in range(probe_steps):
        probe_g.load(*dbatches[i % N_SETS])
        probe_g.replay()
        torch.cuda.synchronize()
        n_k = L.hoc_timer_peek(buf, ids, 8192)
        group_ms += [buf[j] for j in range(n_k) if buf[j] >= 0]
    L.hoc_timer_begin(0)
    # what the one event pair costs, measured live: the step of the bracketed copy against the plain captured step on the
    # same input set, alternating (the two event-record nodes sit at the bracket's ends, so their whole cost -- the extra
    # step time -- is an UPPER bound of what they add to the bracketed interval; the CUPTI timeline shows the same ~4 us)
    cal_plain, cal_inst = [], []
    if world == 1:
        probe_g.load(*dbatches[0])  # (gstep's own input set)
        for _ in range(7):
            cal_plain.append(timed(lambda i: gstep.replay(), 50) / 50)
            cal_inst.append(timed(lambda i: probe_g.replay(), 50) / 50)
    bracket_cost_ms = max(0.0, median(cal_inst) - median(cal_plain)) if cal_plain else None
    del probe_g

    # ---- end-to-end arm: pinned host buffers -> static device buffers -> graph -> host ----
    # A ring of RING captured steps with their own static buffers and their own pinned result slots.  While step i
    # replays, the inputs of steps i+1 .. i+RING-1 are in flight on the copy stream; the host never waits for the step
    # it has just issued: it reads the loss of step i-1 (event wait, one step behind) and goes on issuing.  Every timed
    # step pays exactly one H2D of its inputs and one D2H of its results.
    RING = 3
    hsets = _make_sets(N_SETS, PAIRS, (W, H), None, pin=True, rank=rank, world=world)
    hbatches = [_samples_from_scene(sc) for sc in hsets]
    gsteps = gsteps_dev[:RING] if len(gsteps_dev) >= RING else gsteps_dev + [capture(0) for _ in range(RING - len(gsteps_dev))]
    grad_host = [[torch.empty(PAIRS, 778, 3).pin_memory(), torch.empty(PAIRS, 1502, 3).pin_memory()] for _ in range(RING)]
    loss_host = [torch.empty(()).pin_memory() for _ in range(RING)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    def run_e2e(batches, min_total_ms, frames=None):
        ready = [torch.cuda.Event() for _ in range(RING)]
        done = [torch.cuda.Event() for _ in range(RING)]
        state = {"issued": 0, "loaded": 0, "losses": 0.0}
        for ev in done:
            ev.record(main_stream)

        def prefetch(i):
            slot = i % RING
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[slot])  # the slot's previous step has consumed its inputs
                if frames is None:
                    gsteps[slot].load(*batches[i % N_SETS])
                else:  # decoded uint8 frames: colour jitter + affine crop + normalise + jitter mask on the device
                    fr = frames[i % N_SETS]
                    gsteps[slot].load(*batches[i % N_SETS], skip_images=True)
                    gsteps[slot].load_frames(fr["frames"], fr["affine"], color=fr["color"], orders=fr["orders"])
                ready[slot].record(copy_stream)

        def e2e_step(_):
            i = state["issued"]
            while state["loaded"] < i + RING - 1 + 1:  # keep RING-1 steps of inputs in flight ahead of the replay
                prefetch(state["loaded"])
                state["loaded"] += 1
            slot = i % RING
            main_stream.wait_event(ready[slot])
            loss, gh, go = gsteps[slot].replay()
            grad_host[slot][0].copy_(gh, non_blocking=True)
            grad_host[slot][1].copy_(go, non_blocking=True)
            loss_host[slot].copy_(loss, non_blocking=True)
            done[slot].record(main_stream)
            if i > 0:  # the caller reads the loss of the PREVIOUS step: one step behind, no stall of the one in flight
                done[(i - 1) % RING].synchronize()
                state["losses"] += float(loss_host[(i - 1) % RING])
            state["issued"] = i + 1

        is_img = lambda k: getattr(k, "name", k) in ("IMAGE", "JITTERMASK")
        nbytes = sum(v.numel() * v.element_size() for s_ in batches[0][0] for k, v in s_.items()
                     if torch.is_tensor(v) and not (frames is not None and is_img(k)))
        nbytes += sum(v.numel() * v.element_size() for r in batches[0][1] for v in r.values())
        if frames is not None:
            nbytes += sum(f.numel() for f in frames[0]["frames"]) + PAIRS * (6 * 4 + 3 * 4 + 4 + 8 * 4)
        for i in range(args.warmup):
            e2e_step(i)
        t = timed_blocks(e2e_step, args.steps, min_total_ms=min_total_ms, max_blocks=40)
        torch.cuda.synchronize()
        return median(t), nbytes

    e2e_ms, h2d = run_e2e(hbatches, 500.0)
    d2h = grad_host[0][0].numel() * 4 + grad_host[0][1].numel() * 4 + 4

    # the same with the frames and jitter masks as uint8 in host memory (what an image decoder produces; widened and
    # normalised on the device by hoc_unpack_u8 inside GraphedConsistStep.load): a quarter of the image bytes cross PCIe
    def to_u8(batch):
        samples, results = batch
        out = []
        for s_ in samples:
            d = {}
            for k, v in s_.items():
                name = getattr(k, "name", k)
                if name == "IMAGE":
                    d[k] = ((v + 0.5) * 255.0).round().clamp(0, 255).to(torch.uint8).pin_memory()
                elif name == "JITTERMASK":
                    d[k] = (v * 255.0).round().clamp(0, 255).to(torch.uint8).pin_memory()
                else:
                    d[k] = v
            out.append(d)
        return out, results

    u8batches = [to_u8(bt) for bt in hbatches]
    e2e_u8_ms, h2d_u8 = run_e2e(u8batches, 300.0)

    # and from the DECODED source frames: uint8 [B,270,480,3] frames of FPHAB's quarter size cross PCIe;
    # colour jitter, the shared affine crop / rotation to the input resolution, normalisation and the jitter masks run on
    # the device (inputpipe.augment_frame_pair, bit-compatible with the PIL calls of handobjset.py:336-379)
    from handobjectconsist_b200 import inputpipe
    import numpy as np
    rng = np.random.default_rng(7)
    Hs, Ws = 270, 480
    raw = []
    for k in range(N_SETS):
        affine = np.stack([inputpipe.get_affine_transform((Ws / 2 + rng.uniform(-20, 20), Hs / 2 + rng.uniform(-10, 10)),
                                                          rng.uniform(0.9, 1.2) * Hs, (W, H), rot=rng.uniform(-0.3, 0.3))[0]
                           for _ in range(PAIRS)])
        raw.append({"frames": [torch.randint(0, 256, (PAIRS, Hs, Ws, 3), dtype=torch.uint8).pin_memory() for _ in range(2)],
                    "affine": affine,
                    "color": dict(brightness=rng.uniform(0.5, 1.5, PAIRS), saturation=rng.uniform(0.5, 1.5, PAIRS),
                                  hue=rng.uniform(-0.15, 0.15, PAIRS), contrast=rng.uniform(0.5, 1.5, PAIRS)),
                    "orders": np.stack([[rng.permutation(4) for _ in range(2)] for _ in range(PAIRS)])})
    e2e_raw_ms, h2d_raw = run_e2e(hbatches, 300.0, frames=raw)

    e2e_ms, eager_ms, e2e_u8_ms, e2e_raw_ms = sharding.max_over_ranks([e2e_ms, eager_ms, e2e_u8_ms, e2e_raw_ms], device=dev)
    global_loss = float(sharding.global_mean_loss(gstep.loss))  # the one scalar exchange of the path
    if rank != 0:
        return None

    frames = 2 * PAIRS * world * args.steps
    value = frames / (ms / 1e3)
    peak, peak_src = _peaks()
    Bp, V, Fh, Fo = PAIRS, 2280, 1552, 3000
    Fm = Fh + Fo
    F2 = 2 * Fm                      # faces the rasterizer sees per mesh (fill_back)
    npx, ncrop = Bp * S * S, Bp * H * W
    npx2, nf, nf2 = 2 * npx, Bp * F2, 2 * Bp * F2
    c = coverage
    # ALGORITHMIC bytes per launch: every datum the kernel needs crosses HBM once.  The frame-pair
    # path stacks both renders of a pair along the batch, so the rasterizer kernels run once over 2 * pairs samples.
    algo = {
        # hand / object vertices of both frames + face tables in; faces + vertex values of both renders, the stacked
        # face table and the z-buffer key fill out
        "pair_front": Bp * 2 * V * 12 + (Fh + Bp * Fo) * 24 + nf2 * 72 + 2 * Bp * Fm * 24 + npx2 * 8,
        "raster_zbuf": nf2 * 36 + npx2 * 8,                       # faces in, 8-byte depth/face key per pixel
        # key in; idx4 + alpha4 + rgb12 out; at the covered pixels face + vertex values in, depth4 + weights12 out
        "raster_resolve": npx2 * (8 + 20) + int(c * npx2) * (36 + 36 + 16),
        # finalize + warp, both directions, in the captured step (loss_only: sparse outputs): alpha4 + rgb8 in, valid1 out
        # per pixel; at covered pixels the ignore / occlusion look-ups (idx4 + 2 x (alpha4 + rgb8 + idx4)), the warp's
        # operands (source 12, target 12, jitter 4 + 4) and flow8 + mult4 out (with every output dense, as the eager
        # API hands them out: + flow8 + mult4 + flow_mask2 = 14 B per pixel more)
        "flow_finalize": 2 * ncrop * (12 + 1) + int(c * 2 * ncrop) * (4 + 32 + 32 + 12),
        # warp backward fused with the finalize backward: valid1 in, grad_rgb 12 out per raster pixel; flow8 + mult4 +
        # source 12 + target 12 at the valid pixels
        "warp_photo_bwd": 2 * ncrop * 1 + npx2 * 12 + int(c * 2 * ncrop) * 36,
        # scan pass over both renders with the backward of pair_consist fused in (hoc_raster_bwd_scan_pair_kernel): idx4 +
        # valid1 in, two gradient planes (8) out per pixel; flow8 + mult4 + source 12 + target 12 at the valid pixels;
        # zero-fill of grad_faces (one render) + grad of the vertex values (both) + the scatter's outputs
        "raster_bwd_pixel": (npx2 * 12 + 2 * ncrop * 1 + int(c * 2 * ncrop) * 36 + nf * 36 + nf2 * 36
                             + 2 * 2 * Bp * V * 12),
        # (cover pass: only in configurations whose texture gradient the line pass does not run; not in this workload)
        "raster_bwd_cover": int(c * npx2) * 36 + nf2 * 72,
        "raster_backward": nf * (36 + 12 + 36),                    # depth epilogue (only with dL/ddepth)
        # line pass: the two flow channels of rgb (8), the two gradient planes (8) and idx (4) of the render with the
        # pseudo-gradient once (both axes read the same maps), faces of its covered pixels, grad_faces update; texture
        # gradient of both renders: the second render's gradient planes (8) + weights12 + depth4 at the valid pixels, 9
        # sums per face out
        "raster_bwd_line": npx * (8 + 8 + 4) + nf * (36 + 36) + npx * 8 + int(c * npx2) * 16 + nf2 * 36,
        "mesh_scatter": nf * 36 + nf2 * 36 + 2 * Bp * Fm * 24 + 2 * 2 * Bp * V * 12,
        "pair_back": 2 * 2 * Bp * V * 12 + 2 * Bp * V * 12 + Bp * V * 12,
    }
    kname = {"raster_zbuf": "hoc_raster_zbuf_kernel", "raster_resolve": "hoc_raster_resolve4_kernel",
             "raster_bwd_pixel": "hoc_raster_bwd_scan_pair_kernel",
             "raster_bwd_cover": "hoc_raster_bwd_cover_kernel", "raster_backward": "hoc_raster_bwd_depth_kernel",
             "raster_bwd_line": "hoc_raster_bwd_line_kernel", "warp_photo_bwd": "hoc_warp_photo_pair_backward_kernel",
             "flow_finalize": "hoc_flow_finalize_warp_kernel", "mesh_scatter": "hoc_mesh_scatter_kernel",
             "pair_front": "hoc_pair_front_kernel", "pair_back": "hoc_pair_back_kernel",
             "pair_loss": "hoc_pair_loss_mean_kernel"}
    table = []
    step_ms = ms / args.steps
    for name, v in per_kernel.items():
        avg = sum(v) / len(v)
        ab = algo.get(name)
        table.append({"kernel": name, "launches_per_step": len(v) / probe_steps, "avg_ms": avg,
                      "cupti_ms": None,  # (filled in by the cross-check at the very end of the run)
                      "share_of_step": sum(v) / probe_steps / step_ms,
                      "algorithmic_bytes_per_launch": ab,
                      "achieved_gbs": (ab / (avg * 1e-3) / 1e9) if ab else None})
    table.sort(key=lambda r: -r["share_of_step"])
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "kernel_traffic_r2.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic_tab = json.load(f)
    # `roofline`: the kernel BASELINE.json's north_star names, the rasterizer backward -- three launches here (scan,
    # cover, line pass) over the stacked batch, reported together with the bytes of the backward of both renders counted
    # once: the render whose geometry gradient is needed (H*W*28 + 2F*108 per sample) and the render that only
    # needs its texture gradient (H*W*16 + 2F*72).  `roofline_dominant`: the kernel with the largest share of the step.
    dom = next((r for r in table if r["algorithmic_bytes_per_launch"]), None)
    bwd = [r for r in table if r["kernel"] in ("raster_bwd_pixel", "raster_bwd_cover",
                                               "raster_backward", "raster_bwd_line")]
    bwd_bytes = (npx * 28 + nf * 108) + (npx * 16 + nf * 72)
    bwd_ms_kernels = sum(r["avg_ms"] * r["launches_per_step"] for r in bwd)
    bwd_ms = (sum(group_ms) / len(group_ms)) if group_ms else bwd_ms_kernels
    bwd_traffic = sum(traffic_tab.get(kname[r["kernel"]], 0) for r in bwd) or None
    timing_note = ("CUDA events (external event nodes) around every launch inside an instrumented copy of the captured "
                   "graph; the event nodes add ~3-4 us per kernel, so fractions are slightly pessimistic "
                   "(profiles/timeline_r2.txt holds CUPTI durations of an un-instrumented replay)")

    def roof(r):
        if r is None:
            return None
        return {"kernel": kname.get(r["kernel"], r["kernel"]), "bound": "hbm", "achieved": r["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": r["achieved_gbs"] / peak, "frac_of_nominal_8000": r["achieved_gbs"] / 8000.0,
                "traffic": traffic_tab.get(kname.get(r["kernel"])),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": r["algorithmic_bytes_per_launch"],
                "avg_launch_ms": r["avg_ms"], "launches_timed": int(r["launches_per_step"] * probe_steps),
                "share_of_step": r["share_of_step"], "timing": timing_note}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[{CONFIG}]: {PAIRS} frame pairs/rank (= {2 * PAIRS} mesh renders of configs[1] "
                               f"shape, 9104 faces after fill_back) render->flow->occlusion->warp->masked L1 fwd+bwd, "
                               f"{W}x{H} frames on a {S}x{S} raster, full geometry+texture backward "
                               f"(detach_renders=False), use_backward=True; training step: the visualisation returns of "
                               f"pair_consist (warps / diffs / warp_mask) are not produced",
                   "pairs_per_rank": PAIRS, "image_size": [W, H], "raster_size": S, "faces_per_mesh": F2,
                   "parallelism": f"dp{world}", "measured_coverage": c,
                   "l2": f"{N_SETS} resident input sets rotate (one captured graph each); a step touches ~250 MB > 126 MB L2"},
        "timed_region": {"blocks": len(blocks), "steps_per_block": args.steps, "total_ms": sum(blocks),
                         "block_ms_median": ms, "block_ms_min": min(blocks), "block_ms_max": max(blocks),
                         "note": "the contract's K-step region (barrier + synchronize on both sides, CUDA events, max "
                                 "over ranks), repeated; `value` is the median block"},
        "e2e": {"value": frames / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps, "ring": RING, "numa": numa,
                "note": "GraphedConsistStep.load + replay from pinned fp32 host buffers (the reference's formats), "
                        "results to pinned host memory, loss read one step behind"},
        "e2e_u8": {"value": frames / (e2e_u8_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_u8),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_u8_ms / args.steps,
                   "note": "same call with the frames and jitter masks as uint8 host tensors (widened + normalised on the "
                           "device, hoc_unpack_u8); `e2e` above is the reference's fp32 host format"},
        "e2e_frames": {"value": frames / (e2e_raw_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_raw),
                       "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_raw_ms / args.steps,
                       "note": "the same call fed with DECODED uint8 source frames (480x270): colour jitter, the pair's shared "
                               "affine crop / rotation, normalisation and the jitter masks run on the device "
                               "(GraphedConsistStep.load_frames -> hoc_augment_frame_pair, SURVEY 8f row f3); includes the "
                               "host-side inversion of the affine transforms"},
        "gpu_launches": launches,
        "launches_per_step": launches_per_step,
        "loss_global_mean": global_loss,
        "execution": "forward+backward captured once in a CUDA graph (handobjectconsist_b200.graphed), replayed per step; "
                     "the frame-pair path of consist.py: 9 kernel nodes, no memset / ATen node",
        "with_visuals": ({"value": 2 * PAIRS * args.steps / (vis_ms / 1e3), "ms_per_step": vis_ms / args.steps,
                          "note": "the same captured step when it also produces pair_consist's visualisation returns"}
                         if vis_ms else None),
        "eager": {"value": frames / (eager_ms / 1e3), "ms_per_step": eager_ms / args.steps,
                  "note": "same step with every launch issued from python (CPU-launch-bound)"},
        "clocks": clocks,
        "roofline": {
            "kernel": "rasterizer backward: " + " + ".join(sorted({kname[r["kernel"]] for r in bwd})), "bound": "hbm",
            "achieved": (bwd_bytes / (bwd_ms * 1e-3) / 1e9) if bwd_ms > 0 else None, "peak": peak, "unit": "GB/s",
            "frac": (bwd_bytes / (bwd_ms * 1e-3) / 1e9 / peak) if bwd_ms > 0 else None,
            "frac_of_nominal_8000": (bwd_bytes / (bwd_ms * 1e-3) / 1e9 / 8000.0) if bwd_ms > 0 else None,  # 
            "traffic": bwd_traffic,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": bwd_bytes, "avg_launch_ms": bwd_ms,
            "share_of_step": bwd_ms / step_ms if bwd_ms else None,
            "launches_timed": len(group_ms), "sum_of_per_kernel_brackets_ms": bwd_ms_kernels,
            "cupti": None,  # (filled in by the cross-check at the very end of the run)
            "event_pair": ({"cost_ms_per_step": bracket_cost_ms, "plain_step_ms": median(cal_plain),
                            "bracketed_step_ms": median(cal_inst),
                            "frac_net": ((bwd_bytes / ((bwd_ms - bracket_cost_ms) * 1e-3) / 1e9 / peak)
                                         if bwd_ms > bracket_cost_ms else None),
                            "note": "cost of the bracket's two event-record nodes = step time of the bracketed copy of the "
                                    "graph minus the plain captured step (same inputs, 7 alternating blocks of 50 replays, "
                                    "medians); `frac` above is NOT corrected for it -- `frac_net` is the fraction if all of "
                                    "that cost fell inside the bracket, so the kernels' own fraction lies between the two"}
                           if (bracket_cost_ms is not None and group_ms) else None),
            "timing": "ONE pair of CUDA events (external event nodes of an instrumented copy of the captured graph) around the "
                      "two launches of hoc_pair_backward_raster; the pair adds ~4 us inside the bracket (`cupti` below; "
                      "profiles/timeline_r2.txt holds the CUPTI durations of an un-instrumented replay) and "
                      "`event_pair.cost_ms_per_step` to the step",
            "note": "scan + line pass over the stacked batch of both renders (one launch each); bytes: the render "
                    "with the pseudo-gradient (H*W*28 + 2F*108 per sample) + the texture-only render (H*W*16 + 2F*72).  "
                    "The scan pass also computes the backward of pair_consist (fused in; its own operand bytes are NOT "
                    "added to the numerator), so the fraction is a lower bound for the rasterizer backward alone"},
        "roofline_dominant": roof(dom),
        "kernels": table,
    }
    def cupti_crosscheck(out):
        """Cross-check, NOT a bench value: CUPTI durations (torch.profiler) of the kernels in 20 plain replays of the
        captured step -- what the kernels take without event nodes around them (profiles/timeline.py prints the same as a
        timeline).  Runs after everything else has been measured: the profiler stays loaded in the process."""
        durs = {}
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            for i in range(20):
                graphed_step(i)
            torch.cuda.synchronize()
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA and "hoc_" in ev.name:
                nm = ev.name.split("(")[0].split("<")[0].replace("void ", "").strip()
                durs.setdefault(nm, []).append((ev.time_range.end - ev.time_range.start) * 1e-3)
        for r in out["kernels"]:
            v = durs.get(kname.get(r["kernel"], r["kernel"]))
            r["cupti_ms"] = (sum(v) / len(v)) if v else None
        rows = [r for r in out["kernels"] if r["kernel"] in ("raster_bwd_pixel", "raster_bwd_cover", "raster_backward",
                                                             "raster_bwd_line")]
        if rows and all(r["cupti_ms"] for r in rows):
            t = sum(r["cupti_ms"] * r["launches_per_step"] for r in rows)
            out["roofline"]["cupti"] = {
                "launch_ms": t, "frac": bwd_bytes / (t * 1e-3) / 1e9 / peak,
                "note": "cross-check only (taken under torch.profiler after everything else was measured, so not a bench "
                        "value): sum of the CUPTI durations of the same launches in plain replays of the captured step"}

    out["_cupti_crosscheck"] = cupti_crosscheck  # (main() runs it last and removes the key)
    return out


def run_config4(args, rank, world, local_rank):
    """BASELINE.json configs[3]: the full trainmeshwarp optimisation step under DDP -- ResNet-18 backbone + MLP heads
    (torch / cuDNN, bench_models.py) -> ManoLayer -> geometry head -> WarpRegNet's photometric consistency step (this
    library, one CUDA-graph replay) -> backward -> NCCL all-reduce of the network gradients (DistributedDataParallel,
    overlapped with the backward) -> Adam step (trainmeshwarp.py:181-255, epochpassconsist.py:56-68,
    warpreg.py:81-127).  Weak scaling: PAIRS frame pairs per rank (8 -> global batch 64 on 8 GPUs)."""
    import torch
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    import bench_models
    from handobjectconsist_b200 import _lib, sharding, synth
    from handobjectconsist_b200.queries import BaseQueries, TransQueries
    from handobjectconsist_b200.warpreg import WarpRegNet

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: the product path has no CPU fallback")
    numa = _bind_to_gpu_numa_node(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    W, H, B, hv = WIDTH, HEIGHT, PAIRS, 778
    sf, tf = 1e-4, 100.0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            fn(i)
        t1.record()
        barrier()
        return t0.elapsed_time(t1)

    def median(v):
        v = sorted(v)
        return v[len(v) // 2]

    def timed_blocks(fn, steps, min_total_ms, max_blocks=100):
        first = sharding.max_over_ranks([timed(fn, steps)], device=dev)[0]
        n = int(min(max_blocks, max(1, -(-min_total_ms // max(first, 1e-3)))))
        return sharding.max_over_ranks([first] + [timed(fn, steps) for _ in range(n - 1)], device=dev)

    def batches(device, pin):
        out = []
        for sc in _make_sets(N_SETS, B, (W, H), device, pin=pin, rank=rank, world=world):
            obj_centre = sc["verts1"][:, hv:].mean(1)
            can = (sc["verts1"][:, hv:] - obj_centre[:, None]).contiguous()
            can = can.pin_memory() if (pin and not can.is_cuda) else can
            samples, _ = _samples_from_scene(sc)
            for s_ in samples:
                s_[BaseQueries.OBJCANVERTS] = can
            out.append(({"data": samples, "supervision": ["consist"]}, sc))
        return out

    dbatches = batches(dev, False)
    sc0 = dbatches[0][1]
    torch.manual_seed(0)
    model = bench_models.BenchMeshRegNet(trans_factor=tf, scale_factor=sf).to(dev)
    model.train()
    model.freeze_batchnorm()
    K = sc0["K"]
    f, cc = K[:, 0, 0], K[:, :2, 2]

    def units(centre):  # scale / translation head outputs that put est_c3d at `centre` (inverse of project.py:15-20)
        s_ = (centre[:, 2] - 0.4) / (f * sf)
        t_ = (centre[:, :2] * (f / centre[:, 2])[:, None] - torch.tensor([W / 2.0, H / 2.0], device=dev) + cc) / tf
        return torch.cat([s_[:, None], t_], 1)

    obj_st = torch.cat([units(sc0["verts1"][:, hv:].mean(1)), torch.zeros(B, 3, device=dev)], 1)
    model.set_head_bias(units(sc0["verts1"][:, :hv].mean(1)), obj_st)
    net = WarpRegNet((W, H), model, mano_faces=sc0["faces"][0, :1538].cpu(), use_backward=True, lambda_data=1,
                     lambda_consist=1, progressive_consist=True, progressive_steps=1000, detach_renders=True,
                     graphed=True).to(dev)
    # the synthetic hand is a closed 1552-face template whose own last 14 faces play the wrist cap: use its
    # table instead of MANO's wrist fan, which indexes the real MANO topology
    net.mano_layer.register_buffer("th_faces", sc0["faces"][0, :1552].clone())
    net.step_count = 500  # mid-schedule: both loss terms weighted (warpreg.py:102-110)
    params = [p for p in net.parameters() if p.requires_grad]
    n_param = sum(p.numel() for p in params)
    ddp = DDP(net, device_ids=[local_rank], broadcast_buffers=False, gradient_as_bucket_view=True) if world > 1 else net
    opt = torch.optim.Adam(params, lr=5e-5)
    state = {"loss": None}

    def train_step(batch, sync=True):
        import contextlib
        ctx = ddp.no_sync() if (world > 1 and not sync) else contextlib.nullcontext()
        with ctx:
            loss, agg, _, _ = ddp(batch)
            opt.zero_grad(set_to_none=True)
            loss.backward()
        opt.step()
        state["loss"] = loss.detach()

    launches0 = L.hoc_launch_count(-1)
    train_step(dbatches[0][0])  # captures the consistency step
    for i in range(max(args.warmup, 3)):
        train_step(dbatches[i % N_SETS][0])
    torch.cuda.synchronize()
    l0 = L.hoc_launch_count(-1)
    train_step(dbatches[0][0])
    torch.cuda.synchronize()
    launches_per_step = int(L.hoc_launch_count(-1) - l0)
    graph_nodes = int(l0 - launches0)

    sampler = ClockSampler(local_rank)
    sampler.start()
    blocks = timed_blocks(lambda i: train_step(dbatches[i % N_SETS][0]), args.steps, 1000.0)
    clocks = sampler.finish()
    ms = median(blocks) / args.steps
    nosync_ms = None
    if world > 1:
        nosync_ms = median(timed_blocks(lambda i: train_step(dbatches[i % N_SETS][0], sync=False), args.steps, 300.0)) / args.steps
    # the library's share: the captured consistency step replayed alone
    gstep = net._graph_step
    lib_ms = median(timed_blocks(lambda i: gstep.replay(), args.steps, 200.0)) / args.steps
    # forward + backward of the network alone (no consistency term): data supervision only
    def net_only(i):
        import contextlib
        batch = dict(dbatches[i % N_SETS][0], supervision=["data"])
        with (ddp.no_sync() if world > 1 else contextlib.nullcontext()):
            loss, _, _, _ = ddp(batch)
            opt.zero_grad(set_to_none=True)
            loss.backward()
    net_ms = median(timed_blocks(net_only, args.steps, 300.0)) / args.steps

    # end to end: the batch dicts live in pinned host memory (frames, masks, intrinsics, GT meshes of the reference frame)
    hbatches = batches(None, True)
    h2d = sum(v.numel() * v.element_size() for s_ in hbatches[0][0]["data"] for v in s_.values() if torch.is_tensor(v))
    for i in range(3):
        train_step(hbatches[i % N_SETS][0])

    def e2e_step(i):
        train_step(hbatches[i % N_SETS][0])
        float(state["loss"])  # the training loop logs the loss every step (D2H + sync)

    e2e_ms = median(timed_blocks(e2e_step, args.steps, 500.0)) / args.steps
    loss_val = float(sharding.global_mean_loss(state["loss"]))
    if rank != 0:
        return None
    frames = 2 * B * world
    return {
        "metric": METRIC, "value": frames / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[3]: full trainmeshwarp optimisation step, {B} frame pairs/rank (global {B * world}), "
                               f"{W}x{H}: ResNet-18 (random init, frozen BN) + MLP heads on both frames of every pair -> "
                               f"ManoLayer + geometry head -> WarpRegNet consistency step (CUDA-graph replay of this "
                               f"library, use_backward=True, detach_renders=True like the reference) -> backward -> "
                               f"DDP all-reduce -> Adam",
                   "pairs_per_rank": B, "global_batch": B * world, "image_size": [W, H], "parallelism": f"ddp{world}",
                   "l2": f"{N_SETS} resident batches rotate"},
        "timed_region": {"blocks": len(blocks), "steps_per_block": args.steps, "block_ms_median": median(blocks),
                         "block_ms_min": min(blocks), "block_ms_max": max(blocks)},
        "breakdown_ms": {"step": ms, "network_fwd_bwd_only": net_ms, "consistency_step_library_replay": lib_ms,
                         "library_share_of_step": lib_ms / ms,
                         "step_without_allreduce": nosync_ms,
                         "allreduce_exposed": (ms - nosync_ms) if nosync_ms is not None else None},
        "allreduce": {"params": n_param, "bytes": n_param * 4, "backend": "nccl" if world > 1 else None,
                      "note": "one bucketed all-reduce of the network gradients per step (DistributedDataParallel), "
                              "overlapped with the backward; nothing inside the render / warp kernels communicates"},
        "e2e": {"value": frames / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms, "numa": numa,
                "note": "the same step fed from pinned host batch dicts, loss read back every step"},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "graph_kernel_nodes_at_capture": graph_nodes,
        "loss_global_mean": loss_val, "clocks": clocks,
        "execution": "eager torch network + one CUDA-graph replay of the consistency step (GraphedConsistStep.apply) per "
                     "optimisation step",
    }


def _reference_pairs_per_second(pairs, steps, warmup, threads):
    """Oracle (CPU restatement of the reference algorithm) on `pairs` frame pairs per step."""
    import numpy as np
    import torch

    from handobjectconsist_b200 import synth
    from oracle import pipeline as opipe  # the ONE place bench.py executes oracle/: the CPU baseline

    from oracle import nmr as onmr

    torch.set_num_threads(threads)
    onmr.set_threads(threads)
    times = []
    for i in range(warmup + steps):
        sc = synth.make_scene(pairs, WIDTH, HEIGHT, seed=i)
        t0 = time.perf_counter()
        v1 = sc["verts1"].clone().requires_grad_(True)
        loss, _ = opipe.consist_step(v1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                     sc["jitter_mask_ref"], sc["jitter_mask"], max(WIDTH, HEIGHT), (WIDTH, HEIGHT),
                                     sc["hand_ignore_faces"], detach_renders=False, use_backward=True,
                                     grad_dtype=np.float32)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times), len(times)


def cpu_baseline(sample_pairs=4, steps=4, warmup=1):
    threads = os.cpu_count() or 1
    total, n = _reference_pairs_per_second(sample_pairs, steps, warmup, threads)
    return {"value": 2 * sample_pairs * n / total, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} step(s) of {sample_pairs} frame pairs (of the {PAIRS}-pair workload) at {WIDTH}x{HEIGHT}, "
                      f"oracle/ C restatement (pthreads over pixels/faces) + torch CPU warp/loss"}


def gpu_ref_equiv(steps=3, warmup=1):
    """the reference's launch structure restated on the same GPU (baseline/ref_equiv: per-pixel
    all-faces forward, one serial thread per face for the pseudo-gradient, ATen grid_sample, separate element-wise
    ops), same workload, inputs resident, CUDA-event timed.  A labelled RESTATEMENT -- the upstream CUDA extension is
    not installable here -- used as the denominator of the north star's ">= 10x on one B200"."""
    import torch

    from baseline import ref_equiv  # benchmark baseline; never on the product path

    dev = torch.device("cuda", torch.cuda.current_device())
    sets = _make_sets(2, PAIRS, (WIDTH, HEIGHT), dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def one(i):
        sc = sets[i % 2]
        v1 = sc["verts1"].clone().requires_grad_(True)
        loss, _ = ref_equiv.consist_step(v1, sc["verts2"], sc["faces"], sc["K"], sc["image_ref"], sc["image"],
                                         sc["jitter_mask_ref"], sc["jitter_mask"], max(WIDTH, HEIGHT), (WIDTH, HEIGHT),
                                         sc["hand_ignore_faces"], detach_renders=False, use_backward=True)
        loss.backward()
        return loss

    for i in range(warmup):
        one(i)
    torch.cuda.synchronize()
    t0.record()
    for i in range(steps):
        loss = one(i)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    return {"value": 2 * PAIRS / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "kind": "restated",
            "loss": float(loss.detach()),
            "note": "baseline/ref_equiv: the reference's five rasterizer kernels restated one thread per item "
                    "(all-faces forward, serial per-face pseudo-gradient) + the reference's op-by-op torch composition "
                    "with ATen grid_sample, on this GPU, same workload; NOT the upstream binary (not installable)"}


def geom_head_leg(samples=None, iters=50):
    """the geometry head in front of the path -- ManoAdaptor + centring + recover_3d_proj + both
    projections of the hand (hoc_hand_head_*), Rodrigues + rotation + recover_3d_proj + projection of the object
    (hoc_recover_points_*) -- forward + backward on one hand and one 1502-vertex object per rendered frame, inputs
    resident, CUDA-event timed; beside it the reference's op-by-op ATen composition of the same (restated:
    baseline/ref_equiv/geomhead.py).  Not part of `value`."""
    import torch

    from baseline.ref_equiv import geomhead as ref_head
    from handobjectconsist_b200 import _lib, synth
    from handobjectconsist_b200._geomhead import _HandHeadFunction, _RecoverPointsFunction

    n = 2 * PAIRS if samples is None else samples
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator().manual_seed(0)
    res, sf, tf = (float(SIZE), float(SIZE)), 1e-4, 100.0
    K = synth.camera_intrinsics(n, SIZE, SIZE, dev)
    verts = (torch.randn(n, 778, 3, generator=g) * 0.05).to(dev).requires_grad_(True)
    can = (torch.randn(n, 1502, 3, generator=g) * 0.05).to(dev)
    W = torch.rand(21, 778, generator=g)
    W = (W / W.sum(1, keepdim=True)).to(dev)
    lin = torch.nn.Linear(778, 21, bias=False).to(dev)
    lin.weight.data = W
    lin.weight.requires_grad_(False)
    hst = (torch.randn(n, 3, generator=g) * 0.3).to(dev).requires_grad_(True)
    ost = (torch.randn(n, 6, generator=g) * 0.3).to(dev).requires_grad_(True)
    wv = torch.randn(n, 778, 3, generator=g).to(dev)
    w2 = torch.randn(n, 778, 2, generator=g).to(dev)
    wo = torch.randn(n, 1502, 3, generator=g).to(dev)
    wo2 = torch.randn(n, 1502, 2, generator=g).to(dev)

    def ours():
        out = _HandHeadFunction.apply(verts, None, W, K, hst[:, 0], hst[:, 1:], 9, sf, tf, 0.4, *res)
        obj = _RecoverPointsFunction.apply(can, ost[:, 3:], K, ost[:, 0], ost[:, 1:3], sf, tf, 0.4, *res)
        loss = (out[3] * wv).sum() + (out[5] * w2).sum() + out[4].sum() + (obj[1] * wo).sum() + (obj[2] * wo2).sum()
        torch.autograd.grad(loss, [verts, hst, ost])

    def reference():
        rj, rv, j2, v2 = ref_head.hand_head(verts, lin, 9, K, hst[:, :1], hst[:, 1:], sf, tf, res)
        _, ov, o2 = ref_head.obj_head(can, K, ost[:, :1], ost[:, 1:3], ost[:, 3:], sf, tf, res)
        loss = (rv * wv).sum() + (v2 * w2).sum() + j2.sum() + (ov * wo).sum() + (o2 * wo2).sum()
        torch.autograd.grad(loss, [verts, hst, ost])

    def timed(fn):
        for _ in range(5):
            fn()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(iters):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / iters * 1e3

    us_ours, us_ref = timed(ours), timed(reference)
    L = _lib.lib()
    ids = _lib.KERNEL_IDS
    names = ("hand_head_fwd", "hand_head_bwd", "recover_points_fwd", "recover_points_bwd")
    L.hoc_timer_begin(sum(1 << ids[k] for k in names))
    for _ in range(10):
        ours()
    torch.cuda.synchronize()
    buf, kid = (ctypes.c_float * 64)(), (ctypes.c_int * 64)()
    cnt = L.hoc_timer_end(buf, kid, 64)
    k_us = {}
    for j in range(cnt):
        k_us.setdefault(kid[j], []).append(buf[j] * 1e3)
    return {"samples": n, "fwd_bwd_us": us_ours, "ref_op_by_op_us": us_ref, "speedup": us_ref / us_ours,
            "kernel_us": {k: (sum(k_us[ids[k]]) / len(k_us[ids[k]]) if k_us.get(ids[k]) else None) for k in names},
            "note": "hand head (adaptor, centring, recover_3d_proj, 2 projections) + object head (Rodrigues, rotation, "
                    "recover_3d_proj, projection), forward + backward, eager: 4 launches of this library against the "
                    "reference's op-by-op ATen composition (restated, baseline/ref_equiv/geomhead.py); both launch-bound"}


def head_graph_leg(iters=30):
    """Network outputs -> MANO -> geometry head -> consistency step -> gradients of the network outputs, three ways on
    the bench workload: ONE captured graph (graphed.GraphedHeadConsistStep), the eager front end around the captured
    consistency step (GraphedConsistStep.apply: what a training loop did before the head could be captured), and
    everything eager.  Not part of `value` (the metric is render + warp + photometric from camera-space meshes)."""
    import torch

    from handobjectconsist_b200 import synth, warpbranch
    from handobjectconsist_b200.graphed import GraphedConsistStep, GraphedHeadConsistStep
    from handobjectconsist_b200.mano.manolayer import ManoLayer
    from handobjectconsist_b200.meshregnet import recover_mano_geometry
    from handobjectconsist_b200.neurender.renderer import Renderer
    from handobjectconsist_b200.objbranch import ObjBranch
    from handobjectconsist_b200.optim.pyramidloss import PyramidCriterion
    from handobjectconsist_b200.queries import BaseQueries, TransQueries

    dev = torch.device("cuda", torch.cuda.current_device())
    B, S, hv, sf, tf = PAIRS, SIZE, 778, 1e-4, 100.0
    sc = _make_sets(1, B, S, dev)[0]
    samples, results = _samples_from_scene(sc)
    hand_face, ignore = sc["faces"][0, :1552].clone(), sc["hand_ignore_faces"]
    layer = ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=15, use_pca=True, model=synth.mano_model(seed=3)).to(dev)
    g = torch.Generator().manual_seed(0)
    W = torch.rand(21, hv, generator=g) * (torch.rand(21, hv, generator=g) > 0.9).float()
    W = (W / W.sum(1, keepdim=True)).to(dev)
    K = sc["K"]
    f, cc = K[:, 0, 0], K[:, :2, 2]

    def units(centre):  # scale / translation heads that put est_c3d at `centre` (inverse of project.py:15-20)
        s = (centre[:, 2] - 0.4) / (f * sf)
        t = (centre[:, :2] * (f / centre[:, 2])[:, None] - S / 2.0 + cc) / tf
        return torch.cat([s[:, None], t], 1)

    obj_centre = sc["verts1"][:, hv:].mean(1)
    can = (sc["verts1"][:, hv:] - obj_centre[:, None]).contiguous()
    obj_branch = ObjBranch(trans_factor=tf, scale_factor=sf)
    shape_only = torch.empty(0, 0, S, S)
    inputs = {"pose": (torch.randn(B, 18, generator=g) * 0.4).to(dev), "betas": (torch.randn(B, 10, generator=g) * 0.5).to(dev),
              "hand_st": units(sc["verts1"][:, :hv].mean(1)),
              "obj_st": torch.cat([units(obj_centre), (torch.randn(B, 3, generator=g) * 0.2).to(dev)], 1)}

    def head(inp):
        verts_mm, joints_mm = layer(inp["pose"], th_betas=inp["betas"])
        hand = recover_mano_geometry({"verts3d": verts_mm / 1000, "joints3d": joints_mm / 1000}, K, inp["hand_st"][:, :1],
                                     inp["hand_st"][:, 1:], adaptor=W, mano_center_idx=9, trans_factor=tf,
                                     scale_factor=sf, input_res=(S, S))["recov_handverts3d"]
        sample = {BaseQueries.OBJCANVERTS: can, TransQueries.IMAGE: shape_only, TransQueries.CAMINTR: K}
        return hand, obj_branch(sample, inp["obj_st"])["recov_objverts3d"]

    def renderer():
        return Renderer(image_size=S, R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev),
                        K=torch.ones(1, 3, 3, device=dev), orig_size=S, anti_aliasing=False, fill_back=True, near=0.1,
                        no_light=True)

    crit = PyramidCriterion("l1")
    kw = dict(hand_ignore_faces=ignore, gt_refs=True, first_only=True, use_backward=True, detach_renders=False)
    with torch.no_grad():
        hand0, obj0 = head(inputs)
    example = [{"recov_handverts3d": hand0, "recov_objverts3d": obj0}, results[1]]
    full = GraphedHeadConsistStep(head, inputs, renderer(), crit, (S, S), hand_face, samples, example, warmup=1, **kw)
    part = GraphedConsistStep(renderer(), crit, (S, S), hand_face, samples, example, warmup=1, **kw)
    rend = renderer()

    def leaves():
        return {k: v.detach().requires_grad_(True) for k, v in inputs.items()}

    def eager_head_graphed_step():
        inp = leaves()
        hand, obj = head(inp)
        part.apply(samples, [{"recov_handverts3d": hand, "recov_objverts3d": obj}, results[1]]).backward()

    def all_eager():
        inp = leaves()
        hand, obj = head(inp)
        loss, _ = warpbranch.forward(samples, [{"recov_handverts3d": hand, "recov_objverts3d": obj}, results[1]],
                                     hand_face, rend, (S, S), crit, **kw)
        loss.backward()

    def timed(fn):
        for _ in range(5):
            fn()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(iters):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / iters

    ms_full = timed(full.graph.replay)
    ms_part = timed(eager_head_graphed_step)
    ms_eager = timed(all_eager)
    return {"pairs": B, "graphed_with_head_ms": ms_full, "eager_head_plus_graphed_step_ms": ms_part,
            "all_eager_ms": ms_eager, "frames_per_s_graphed_with_head": 2 * B / (ms_full / 1e3),
            "loss": float(full.loss),
            "note": "network outputs (pose, shape, scale / translation / rotation heads) -> ManoLayer -> ManoAdaptor + "
                    "recover_3d_proj -> ObjBranch -> consistency step -> gradients of the network outputs; "
                    "GraphedHeadConsistStep replays all of it as one CUDA graph"}


def mano_leg(hands=None, iters=50):
    """ManoLayer forward + backward (hoc_mano_forward / hoc_mano_backward) on one hand per rendered
    frame of the workload, inputs resident, CUDA-event timed.  Not part of `value` (the metric is render + warp +
    photometric); reported beside it because the skinning is on the path."""
    import torch

    from handobjectconsist_b200 import synth
    from handobjectconsist_b200.mano.manolayer import ManoLayer

    hands = 2 * PAIRS if hands is None else hands
    dev = torch.device("cuda", torch.cuda.current_device())
    layer = ManoLayer(center_idx=9, flat_hand_mean=False, ncomps=15, side="right", use_pca=True,
                      model=synth.mano_model(seed=3)).to(dev)
    g = torch.Generator().manual_seed(0)
    pose = (torch.randn(hands, 18, generator=g) * 0.68).to(dev).requires_grad_(True)
    betas = (torch.randn(hands, 10, generator=g) * 0.02).to(dev).requires_grad_(True)
    gv = torch.randn(hands, 778, 3, generator=g).to(dev)

    def one():
        verts, joints = layer(pose, th_betas=betas)
        torch.autograd.grad((verts * gv).sum() + joints.sum(), [pose, betas])

    for _ in range(5):
        one()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(iters):
        one()
    t1.record()
    torch.cuda.synchronize()
    us = t0.elapsed_time(t1) / iters * 1e3
    # the two kernels alone (event-bracketed by the library)
    from handobjectconsist_b200 import _lib
    L = _lib.lib()
    ids = _lib.KERNEL_IDS
    L.hoc_timer_begin((1 << ids["mano_fwd"]) | (1 << ids["mano_bwd"]))
    for _ in range(10):
        one()
    torch.cuda.synchronize()
    buf, kid = (ctypes.c_float * 64)(), (ctypes.c_int * 64)()
    n = L.hoc_timer_end(buf, kid, 64)
    k_us = {}
    for j in range(n):
        k_us.setdefault(kid[j], []).append(buf[j] * 1e3)
    per_iter = lambda v: (sum(v) / 10.0) if v else None  # 10 timed iterations; the backward is two launches
    return {"hands": hands, "fwd_bwd_us": us, "hands_per_s": hands / (us * 1e-6),
            "forward_kernel_us": per_iter(k_us.get(ids["mano_fwd"], [])),
            "backward_kernels_us": per_iter(k_us.get(ids["mano_bwd"], [])),
            "note": "ManoLayer forward + backward, eager (launch-bound: two kernels of this library plus the torch ops "
                    "of the toy loss), synthetic MANO-shaped model"}


def run_reference(args, rank):
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    pairs = 4
    steps = max(1, min(args.steps, 6))
    warmup = min(args.warmup, 1)
    total, n = _reference_pairs_per_second(pairs, steps, warmup, threads)
    value = 2 * pairs * n / total
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
        "warmup": warmup, "ms_per_step": total / n * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[{CONFIG}] on host cores: bounded sample of {pairs} frame pairs/step at "
                               f"{WIDTH}x{HEIGHT} (same scene generator, full backward, use_backward=True)",
                   "pairs_per_step": pairs, "image_size": SIZE if WIDTH == HEIGHT else [WIDTH, HEIGHT],
                   "faces_per_mesh": 2 * 4552},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} timed step(s) of {pairs} frame pairs; the reference has no CPU renderer and "
                                   f"its CUDA extension is absent, so this is the oracle/ restatement"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries ONE JSON line: keep a private handle on the real stdout and point file descriptor 1 at stderr,
    so that nothing a library prints (NCCL's version banner, warnings) can land next to the result."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(obj):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-equiv", action="store_true", help="skip the GPU reference-equivalent baseline leg")
    ap.add_argument("--eager-only", action="store_true", help="profiling aid: run only the eager arm (ncu)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5],
                    help="BASELINE.json workload: 2 (default, the metric's own: 16 pairs/rank at 256x256), 4 (configs[3]: "
                         "full trainmeshwarp step under DDP, 8 pairs/rank), 5 (configs[4]: 32 pairs/rank at 480x270)")
    ap.add_argument("--pairs", type=int, default=None, help="frame pairs per rank and step (default: the config's)")
    ap.add_argument("--size", type=int, default=None, help="square frame / raster side (default: the config's)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    args = ap.parse_args()
    _claim_stdout()
    pairs, width, height = {2: (16, 256, 256), 4: (8, 256, 256), 5: (32, 480, 270)}[args.config]
    if args.size is not None:
        width = height = args.size
    width = args.width if args.width is not None else width
    height = args.height if args.height is not None else height
    g = globals()
    g["PAIRS"], g["WIDTH"], g["HEIGHT"], g["CONFIG"] = (args.pairs if args.pairs is not None else pairs), width, height, args.config
    g["SIZE"] = max(width, height)
    if args.config == 5:
        g["METRIC"] = "render+warp+photometric fwd+bwd frames/sec @480x270 (480x480 raster)"
    elif args.config == 4:
        g["METRIC"] = "full trainmeshwarp step (ResNet-18 + MANO + render + warp) frames/sec @256x256"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        out = run_reference(args, rank)
        if out is not None:
            _emit(out)
        return

    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        import torch
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    args.warmup = max(args.warmup, 3)
    if args.config == 4:
        out = run_config4(args, rank, world, local_rank)
        if out is not None:
            _emit(out)
        if world > 1:
            dist.destroy_process_group()
        return
    out = run_native(args, rank, world, local_rank)
    legs = args.config == 2 and WIDTH == HEIGHT
    if out is not None:
        if world == 1 and not args.no_ref_equiv and legs:
            try:
                out["gpu_ref_equiv"] = gpu_ref_equiv()
                out["gpu_ref_equiv"]["speedup_of_value"] = out["value"] / out["gpu_ref_equiv"]["value"]
            except Exception as exc:  # a baseline leg must never cost the bench line
                out["gpu_ref_equiv"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
        if world == 1 and legs:
            try:
                out["mano"] = mano_leg()
            except Exception as exc:
                out["mano"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
            try:
                out["geom_head"] = geom_head_leg()
            except Exception as exc:
                out["geom_head"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
            try:
                out["head_graph"] = head_graph_leg()
            except Exception as exc:
                out["head_graph"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline()
        crosscheck = out.pop("_cupti_crosscheck", None)
        if crosscheck is not None and world == 1:
            try:
                crosscheck(out)
            except Exception as exc:  # a cross-check must never cost the bench line
                out["roofline"]["cupti"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
        _emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
