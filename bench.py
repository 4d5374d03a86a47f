#!/usr/bin/env python
"""bench.py -- particles scored/s (and tracked frames/s per camera) of the appearance-scoring path.

Workload (BASELINE.json configs[1]): one camera per GPU, 8 objects, HaarAndColorHistogram features
(250 Haar + 512 MDCH -> 152 selected), WeightedStumps + MILBoost, 2000 particles/object, 640x480 synthetic frames.

  step   = one frame of the SCORE PATH for all 8 objects: integral image + MDCH bin image + predict + test set +
           colour-histogram rows + fused Haar/stump response + likelihood map + weight update + normalise +
           N_eff + systematic resample + read-out  (SURVEY.md 8d "particles scored/s").
  value  = whole-job particles scored / s with frames and prediction noise already resident in HBM
           (frame pool larger than L2 is cycled), timed with CUDA events on the launching stream.
  e2e    = the same metric through the reference-facing C++ host API (ParticleFilterTracker::TrackCameraObjects)
           with HOST buffers: per step the H2D copy of the 4 frame planes + noise from pinned memory and the D2H
           read-out of the per-object summaries are inside the timed region.
  extra  = tracked_frames_per_sec_per_camera: full frames (score path + UpdateClassifier) through the host API; the
           default-strategy training sets are generated on the device behind the frame's scoring kernels and the
           trackers' RNG streams are device-resident meanwhile (MCT_HOST_TRAINING_SETS=1: host enumeration).
  --config 0..4 = the other BASELINE configurations (0: 1 object, Haar only -- only the gray plane is handed over;
           2 / 3: the camera network with the geometric / appearance fuser, one camera per rank under torchrun;
           4: the Haar gather sweep).  MCT_E2E_STREAM=1: the streaming form of the per-frame call for the e2e leg.

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads via the reference's own
omp pragmas) on the same config and metric.  Multi-GPU: one process per GPU (torchrun), one camera per rank,
no data-path collective (configs[1] has no fuser) -> weak scaling.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 640, 480
N_OBJ, N_PART = 8, 2000
N_HAAR, N_BINS, PCT = 250, 8, 20
FEAT, WEAK, STRONG = 3, 1, 2          # HAAR_COLOR_HISTOGRAM, WEIGHTED_STUMP, MILBoost
SD = (20.0, 20.0, 1e-7, 1e-7)
POOL = 128                            # frames in the resident pool: 128 * 4 planes * 300 KB = 157 MB > 126 MB L2
METRIC = "particles_scored_per_sec"
CONFIG = {"workload": "configs[1]: 1 camera/GPU, 8 objects, Haar(250)+MDCH(512) -> 152 selected, WeightedStumps+MILBoost, "
                      "2000 particles/object, 640x480 synthetic",
          "frame": [W, H], "objects_per_camera": N_OBJ, "particles_per_object": N_PART, "features": N_HAAR + N_BINS ** 3,
          "selected": int((N_HAAR + N_BINS ** 3) * PCT / 100),
          "l2_policy": "inputs larger than L2: %d-frame device-resident pool (157 MB) cycled" % POOL}


# BASELINE.json configs[0..4].  0 / 1: one camera per GPU without a fuser (replicas, no data-path collective); 2 / 3: the camera
# network with a fuser (one camera per rank under torchrun, or all cameras on this GPU when run alone); 4: Haar scoring sweep.
CONFIGS = {
    0: dict(W=640, H=480, n_obj=1, n_part=1000, n_haar=250, feat=0, weak=0, strong=2,
            workload="configs[0]: 1 camera, 1 object, Haar(250) -> 50 selected, OnlineStumps+MILBoost, 1000 particles, 640x480 synthetic"),
    1: dict(W=640, H=480, n_obj=8, n_part=2000, n_haar=250, feat=3, weak=1, strong=2, workload=CONFIG["workload"]),
    2: dict(W=640, H=480, n_obj=8, n_part=1000, n_haar=250, feat=0, weak=0, strong=2, cameras=4, camouflage=None,
            net=dict(geometric_fusion=1),
            workload="configs[2]: 4 cameras x 8 objects, Haar(250) -> 50 selected, OnlineStumps+MILBoost, 1000 particles/object, 640x480, "
                     "GeometryBasedInformationFuser (principal-axis intersection + Kalman) with feedback, one camera per GPU"),
    3: dict(W=1280, H=720, n_obj=16, n_part=1000, n_haar=0, feat=2, weak=0, strong=4, cameras=8, camouflage=30,
            net=dict(appearance_fusion=2, af_refresh=1),
            workload="configs[3]: 8 cameras x 16 objects, 1280x720, MDCH(512) -> 102 selected, OnlineStumps+MILAnyBoost, 1000 particles/object, "
                     "AppearanceBasedInformationFuser (MILBoost over MDCH, refresh every frame), one camera per GPU"),
}


def configure(k):
    global W, H, N_OBJ, N_PART, N_HAAR, FEAT, WEAK, STRONG, CONFIG
    c = CONFIGS[k]
    W, H, N_OBJ, N_PART, N_HAAR, FEAT, WEAK, STRONG = c["W"], c["H"], c["n_obj"], c["n_part"], c["n_haar"], c["feat"], c["weak"], c["strong"]
    F = {0: N_HAAR, 2: N_BINS ** 3, 3: N_HAAR + N_BINS ** 3}[FEAT]
    CONFIG = {"workload": c["workload"], "frame": [W, H], "objects_per_camera": N_OBJ, "particles_per_object": N_PART, "features": F,
              "selected": int(F * PCT / 100),
              "l2_policy": "inputs larger than L2: %d-frame device-resident pool (%d MB) cycled" % (POOL, POOL * 4 * W * H // 1000000)}


def pingpong(i, n):
    """frame index sequence 1,2,..,n-1,n-2,..,1,2.. so that object motion stays continuous"""
    period = 2 * (n - 2)
    k = i % period
    return 1 + (k if k < n - 1 else period - k)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md).  NVML in-process (a query takes tens of
    microseconds, so even a timed region of a few milliseconds collects samples); nvidia-smi as the fallback."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices; NVML enumerates the physical ones
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            nv = self.nvml
            names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
            while True:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    self.samples.append([str(sm), str(self.max_sm)] + ["Active" if mask & bit else "Not Active" for _, bit in names])
                except Exception:
                    pass
                if self.stop_flag:
                    return
                time.sleep(0.05)         # 20 samples per second: no measurable load on the host cores the ranks share
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while True:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            if self.stop_flag:
                return
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=10)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max(int(s[1]) for s in self.samples if s[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 2 + k and s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_sequence(camera):
    from multicameratracking_b200 import synth
    seq = synth.Sequence(W, H, N_OBJ, camera=camera, n_frames=POOL)
    frames = [seq.frame(t) for t in range(POOL)]
    return seq, frames


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_setup(orc, seq, frames):
    from oracle import oracle_py
    f0 = orc.frame(*frames[0])
    F = CONFIG["features"]
    K = int(F * PCT / 100)
    trackers = []
    for o in range(N_OBJ):
        rng = orc.rng(o + 1)
        t = orc.haar_generate(rng, N_HAAR, seq.w0, seq.h0) if N_HAAR > 0 else None
        clf = orc.clf_create(FEAT, N_HAAR, N_BINS, seq.w0, seq.h0, WEAK, STRONG, K, 0.85, t)
        tp = oracle_py.TrackerParams(N_PART, SD[0], SD[1], SD[2], SD[3], 10.0, 7.0, 30, 30, 100000, 20, 0, 0, 0, 30)
        init = np.array(seq.box(0, o), np.float32)
        trk = orc.lib.orc_tracker_create(C.byref(tp), clf, C.byref(f0), init.ctypes.data_as(C.POINTER(C.c_float)), C.byref(rng))
        trackers.append((trk, clf, rng))
    return trackers


def oracle_score_frame(orc, trackers, frames, t, objs, phases=1):
    """one frame of the score path (incl. the integral image) for `objs` objects; returns particles scored"""
    from multicameratracking_b200 import synth
    g, r, gg, b = frames[t]
    of = orc.frame(g, r, gg, b)               # computes the integral image (Matrixu::initII)
    box = np.zeros(4, np.int32)
    scored = 0
    for o in objs:
        nz = synth.noise(N_PART, SD, obj=o, t=t)
        orc.lib.orc_tracker_track(trackers[o][0], C.byref(of), nz.ctypes.data_as(C.POINTER(C.c_float)),
                                  box.ctypes.data_as(C.POINTER(C.c_int)), phases)
        scored += orc.lib.orc_tracker_last_valid(trackers[o][0])
    return scored


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle_py import Oracle
    cores = os.cpu_count() or 1
    try:
        orc = Oracle(omp=True)
        orc.lib.orc_clf_set_threads(cores)
        kind_threads = cores
    except OSError:
        orc = Oracle(omp=False)
        kind_threads = 1
    seq, frames = make_sequence(0)
    trackers = oracle_setup(orc, seq, frames)
    objs = list(range(N_OBJ))
    # bounded sample: shrink the number of objects per step if a full frame would blow the time budget
    t0 = time.perf_counter()
    oracle_score_frame(orc, trackers, frames, 1, objs[:1])
    per_obj = time.perf_counter() - t0
    budget = 150.0
    n_o = max(1, min(N_OBJ, int(budget / max(per_obj * (args.steps + args.warmup), 1e-9))))
    objs = objs[:n_o]
    for i in range(args.warmup):
        oracle_score_frame(orc, trackers, frames, pingpong(i, POOL), objs)
    scored = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        scored += oracle_score_frame(orc, trackers, frames, pingpong(args.warmup + i, POOL), objs)
    dt = time.perf_counter() - t0
    val = scored / dt
    sample = "%d steps x %d of %d objects x %d particles (score path incl. integral image)" % (args.steps, n_o, N_OBJ, N_PART)
    ref_build = reference_build_leg(seq, frames)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "particles/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {"value": val, "unit": "particles/s", "cores": kind_threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_build": ref_build}))


def reference_build_leg(seq, frames):
    """The reference's OWN code (oracle/_ref = /root/reference/src compiled against header shims), single thread as it ships
    (obj/Makefile has no -fopenmp), on a bounded sample: one object, score path only.  Reported next to the OpenMP port above,
    which is the stronger baseline and stays the line's value."""
    try:
        from multicameratracking_b200 import synth
        from oracle.ref_py import Ref, TrackerParams
        if not Ref.available():
            return {"unavailable": "oracle/_ref not built and /root/reference absent"}
        ref = Ref()
        weak = 0 if WEAK == 1 else WEAK          # WeightedStumps dereferences NULL weight lists in the reference
        fr = ref.frame(*frames[0])
        ref.seed(1)
        tp = TrackerParams(N_PART, SD[0], SD[1], SD[2], SD[3], 10.0, 7.0, 30, 30, 100000, 20, 0, 0, 0, 30)
        trk = ref.tracker_create(True, tp, FEAT, max(N_HAAR, 1), N_BINS, weak, STRONG, CONFIG["selected"], 0.85, fr,
                                 np.array(seq.box(0, 0), np.float32), 64)
        n, scored, t0 = 0, 0, time.perf_counter()
        while time.perf_counter() - t0 < 10.0 and n < 20:
            f2 = ref.frame(*frames[pingpong(n, POOL)])
            ref.tracker_track(trk, f2, synth.noise(N_PART, SD, obj=0, t=n + 1), 1)
            scored += ref.lib.ref_tracker_test_set_size(trk[0])
            ref.frame_free(f2)
            n += 1
        dt = time.perf_counter() - t0
        return {"value": scored / dt, "unit": "particles/s", "cores": 1, "kind": "reference",
                "sample": "%d frames x 1 of %d objects x %d particles, score path incl. initII%s" % (
                    n, N_OBJ, N_PART, "; OnlineStumps in place of WeightedStumps (the reference crashes on its NULL weight lists)" if WEAK == 1 else "")}
    except Exception as e:                      # noqa: BLE001
        return {"unavailable": repr(e)}


# ------------------------------------------------------------------------------------------------ our arm
def cpu_baseline_leg(phases=1):
    """oracle (single thread, as the reference ships) on a bounded sample: a few frames of 2 objects.  phases=3: full frames
    (score path + UpdateClassifier), reported as tracked frames/s per camera (the second metric of BASELINE.json)"""
    from oracle.oracle_py import Oracle
    orc = Oracle(omp=False)
    seq, frames = make_sequence(0)
    trackers = oracle_setup(orc, seq, frames)
    objs = [0, 1][:N_OBJ]
    oracle_score_frame(orc, trackers, frames, 1, objs)
    scored, t0, n = 0, time.perf_counter(), 0
    while time.perf_counter() - t0 < (12.0 if phases == 1 else 8.0) and n < 40:
        scored += oracle_score_frame(orc, trackers, frames, pingpong(n + 1, POOL), objs, phases)
        n += 1
    dt = time.perf_counter() - t0
    if phases != 1:
        # frames/s for a whole camera = measured object-frames/s divided by the objects per camera
        return {"value": n * len(objs) / float(N_OBJ) / dt, "unit": "tracked frames/s per camera", "cores": 1, "kind": "port",
                "sample": "%d full frames (score + UpdateClassifier) x %d of %d objects x %d particles, single thread" % (n, len(objs), N_OBJ, N_PART)}
    return {"value": scored / dt, "unit": "particles/s", "cores": 1, "kind": "port",
            "sample": "%d frames x %d of %d objects x %d particles, single thread (reference builds compile omp out)" % (n, len(objs), N_OBJ, N_PART)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from multicameratracking_b200 import capi, synth
    from multicameratracking_b200 import host as mhost

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    # a dedicated (non-default) torch stream: the C-ABI launches on it and torch.cuda.Event times it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    seq, frames = make_sequence(rank)
    cam = mhost.Camera(local, n_frames=POOL + 8, stream=stream.cuda_stream)
    L = capi.load()
    ctx_h = C.c_void_p(cam.L.mcth_camera_ctx(cam.h))
    cam.set_frame(*frames[0])
    for o in range(N_OBJ):
        cam.add_object(seq.box(0, o), feature_type=FEAT, n_haar=N_HAAR, n_bins=N_BINS, weak_type=WEAK, strong_type=STRONG,
                       percent_selected=PCT, n_particles=N_PART, sd_x=SD[0], sd_y=SD[1], sd_sx=SD[2], sd_sy=SD[3], rng_seed=o + 1)
    cam.init_trackers()
    cam.sync()

    total = args.steps + args.warmup
    # resident inputs: frame pool + per-step noise on the device; pinned copies for the e2e leg
    planes_dev = [[torch.from_numpy(p).to(dev) for p in fr] for fr in frames]
    planes_pin = [[torch.from_numpy(p).pin_memory() for p in fr] for fr in frames]
    # the frame as the reference's Camera gets it from the video (packed BGR, Camera.cpp:449-494): the e2e leg uploads this
    bgr_pin = [torch.from_numpy(np.ascontiguousarray(np.stack([fr[3], fr[2], fr[1]], axis=2))).pin_memory() for fr in frames]
    noise_np = [np.stack([synth.noise(N_PART, SD, camera=rank, obj=o, t=i) for o in range(N_OBJ)]) for i in range(total)]
    noise_dev = [torch.from_numpy(n).to(dev) for n in noise_np]
    noise_pin = [torch.from_numpy(n).pin_memory() for n in noise_np]
    torch.cuda.synchronize()

    vp = C.c_void_p

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        t = pingpong(i, POOL)
        g, r, gg, b = planes_dev[t]
        cam.set_frame_device(W, H, W, g.data_ptr(), r.data_ptr(), gg.data_ptr(), b.data_ptr(), ready=True)   # static resident pool
        nz = noise_dev[i % total]
        rc = cam.L.mcth_camera_track_resident(cam.h, t, vp(nz.data_ptr()))
        if rc:
            cam._chk(rc)

    host_t = [0, 0, 0, 0, 0]      # ns spent in set_frame / prefetch_frame / prefetch_noise / track, and calls (MCT_BENCH_HOST_TIMING=1)
    host_timing = os.environ.get("MCT_BENCH_HOST_TIMING") == "1"

    # ctypes views of the (fixed) pinned buffers, made once: a C++ caller would pass the pointers directly
    # Haar-only classifiers read the gray plane only (FeatureParameters.cpp:16-20): the colour planes are then not handed to the
    # library at all (NULL planes are part of the interface), so they are neither uploaded nor binned
    gray_only = FEAT == 0 and not args.e2e_bgr
    planes_ptr = None if args.e2e_bgr else [tuple(vp(x.data_ptr()) if (k == 0 or not gray_only) else None for k, x in enumerate(pl)) for pl in planes_pin]
    bgr_ptr = [vp(x.data_ptr()) for x in bgr_pin] if args.e2e_bgr else None
    noise_ptr = [vp(x.data_ptr()) for x in noise_pin]
    noise_len = [x.numel() for x in noise_pin]
    out_box = np.zeros((N_OBJ, 4), np.int32)
    out_ptr = out_box.ctypes.data_as(C.POINTER(C.c_int))
    Lh, camh = cam.L, cam.h

    # one call per video frame (mcth_camera_step: make frame t current, announce frame t+1 and its noise, track)
    plane_arr = None if args.e2e_bgr else [(vp * 4)(*pl) for pl in planes_ptr]

    def step_e2e(i, phases, last=False):
        t = pingpong(i, POOL)
        if plane_arr is not None and not host_timing:
            j = (i + 1) % total
            if last:                      # closes a streaming run: nothing announced, nothing queued ahead
                cam._chk(Lh.mcth_camera_step(camh, t, plane_arr[t], None, W, H, W, noise_ptr[i % total], None, 0, phases, out_ptr))
                return out_box
            cam._chk(Lh.mcth_camera_step(camh, t, plane_arr[t], plane_arr[pingpong(i + 1, POOL)], W, H, W, noise_ptr[i % total], noise_ptr[j],
                                         noise_len[j], phases, out_ptr))
            return out_box
        t0 = time.perf_counter_ns() if host_timing else 0
        if planes_ptr is not None:
            g, r, gg, b = planes_ptr[t]
            cam._chk(Lh.mcth_camera_set_frame(camh, g, r, gg, b, W, H, W, 0))
            t1 = time.perf_counter_ns() if host_timing else 0
            # the upload of the NEXT frame (pinned host memory) is started now and overlaps this frame's kernels, as a video
            # reader would decode ahead (Matrixu::PrefetchFrame); one frame is still copied per step
            g2, r2, gg2, b2 = planes_ptr[pingpong(i + 1, POOL)]
            cam._chk(Lh.mcth_camera_prefetch_frame(camh, g2, r2, gg2, b2, W, H, W))
        else:
            cam._chk(Lh.mcth_camera_set_frame_bgr(camh, bgr_ptr[t], W, H, 3 * W, 0, 0))
            t1 = time.perf_counter_ns() if host_timing else 0
            cam._chk(Lh.mcth_camera_prefetch_frame_bgr(camh, bgr_ptr[pingpong(i + 1, POOL)], W, H, 3 * W, 0))
        t2 = time.perf_counter_ns() if host_timing else 0
        # the next frame's Gaussian rows ride along as well (mct_noise_prefetch)
        cam._chk(Lh.mcth_camera_prefetch_noise(camh, noise_ptr[(i + 1) % total], noise_len[(i + 1) % total]))
        t3 = time.perf_counter_ns() if host_timing else 0
        cam._chk(Lh.mcth_camera_track(camh, t, noise_ptr[i % total], phases, out_ptr))
        if host_timing:
            t4 = time.perf_counter_ns()
            host_t[0] += t1 - t0; host_t[1] += t2 - t1; host_t[2] += t3 - t2; host_t[3] += t4 - t3; host_t[4] += 1
        return out_box

    # The legs below run on the same trackers one after the other, so the frame counter keeps running across them (the
    # synthetic objects move continuously through the ping-pong pool; a leg that restarted at frame 0 would find the objects
    # far from where the previous leg left the trackers, and lost trackers spread their particles over the slow paths).
    clock = [0]

    def timed(fn, n_warm, n_steps, closer=None):
        offset = clock[0]
        clock[0] += n_warm + n_steps + (1 if closer else 0)
        for i in range(n_warm):
            fn(offset + i)
        barrier()
        capi.scored_particles(L, ctx_h, reset=True)
        l0 = cam.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(n_steps):
            fn(offset + n_warm + i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        res = ms, capi.scored_particles(L, ctx_h), cam.launches - l0
        if closer:
            closer(offset + n_warm + n_steps)       # (untimed: consumes the frame a streaming leg queued ahead)
            barrier()
        return res

    # ---- 1. full frames (score + UpdateClassifier) through the host API, first: the trackers start on their objects ----
    ms_full, _, _ = timed(lambda i: step_e2e(i, 3), args.warmup, args.steps)
    # per-kernel CUDA-event times of full frames (the UpdateClassifier kernels ride after the score path)
    L.mct_profile_enable(ctx_h, 1)
    n_pf = min(args.steps, 20)
    for i in range(n_pf):
        step_e2e(clock[0] + i, 3)
    clock[0] += n_pf
    prof_full = capi.profile_read(L, ctx_h)
    L.mct_profile_enable(ctx_h, 0)
    host_t[:] = [0, 0, 0, 0, 0]
    # ---- 2. resident-input throughput of the score path (value) ----
    sampler = ClockSampler(local)
    sampler.start()
    ms, scored, launches = timed(step_resident, args.warmup, args.steps)
    sampler.stop_flag = True
    # ---- 3. per-kernel CUDA-event timing of the same steps (roofline of the dominant kernel) ----
    # (every step is drained before the next one is issued: in the free-running leg above the host runs ahead and the preparation
    #  of frame t+1 shares the SMs with the kernels of frame t, which would be charged to whichever kernel it delays)
    L.mct_profile_enable(ctx_h, 1)
    for i in range(args.steps):
        step_resident(clock[0] + i)
        cam.sync()
    clock[0] += args.steps
    prof = capi.profile_read(L, ctx_h)
    L.mct_profile_enable(ctx_h, 0)
    # ---- 4. end to end through the host API, host buffers ----
    # streaming calls (phases 1 | 16): a call returns frame t's boxes after it has queued frame t + 1, whose planes and noise it uploads;
    # every timed call still carries one frame's host -> device copies, one frame of kernels and one frame's read-out (the frame
    # queued by the last timed call is consumed by an untimed closing call).  Opt-in (MCT_E2E_STREAM=1): on B200 the synchronous
    # call per frame measures 134.5 us against 138.3 us streaming, so the default stays the reference's
    # synchronous per-frame interface.
    stream_e2e = plane_arr is not None and not host_timing and bool(os.environ.get("MCT_E2E_STREAM"))
    if stream_e2e:
        ms_e2e, scored_e2e, _ = timed(lambda i: step_e2e(i, 1 | 16), args.warmup, args.steps, closer=lambda i: step_e2e(i, 1 | 16, last=True))
    else:
        ms_e2e, scored_e2e, _ = timed(lambda i: step_e2e(i, 1), args.warmup, args.steps)
    if host_timing:
        sys.stderr.write("host us/step (score e2e): set_frame %.1f prefetch_frame %.1f prefetch_noise %.1f track %.1f\n" % tuple(1e-3 * x / max(1, host_t[4]) for x in host_t[:4]))

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- 5. (N > 1) the cross-camera exchange of the geometric fuser: device-side pack + one NCCL allgather per frame
    fuser = None
    if world > 1:
        from multicameratracking_b200.network import CameraNetworkExchange
        ex = CameraNetworkExchange(N_OBJ)
        pts = torch.zeros((N_OBJ, 2), dtype=torch.float32, device=dev)
        for _ in range(5):
            cam.pack_foot_points(pts.data_ptr()); allp = ex.allgather_foot_points(pts)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(50):
            cam.pack_foot_points(pts.data_ptr()); allp = ex.allgather_foot_points(pts)
        f1.record(stream)
        barrier()
        ok = bool(torch.equal(allp[rank], pts)) and tuple(allp.shape) == (world, N_OBJ, 2)
        fuser = {"allgather_us_per_frame": allmax(1e3 * f0.elapsed_time(f1) / 50), "payload_bytes_per_rank": N_OBJ * 8, "verified": ok,
                 "what": "mct_fuser_pack_foot_points + torch.distributed.all_gather_into_tensor (NCCL) per frame"}
        # the same payload through the library's own peer-memory kernels: the foot-point kernel stores into every camera's
        # receive buffer over NVLink, a bounded wait gathers the frame (host wall clock: the call returns the gathered points).
        # Every step that can fail locally is followed by an all-reduce of a success flag, so that the ranks never wait for
        # each other inside a collective one of them did not reach; the measurement is skipped (and says why) otherwise.
        def all_ok(flag):
            t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(t.item() > 0.5)
        try:
            cctx = cam.capi_context()
            hbuf = C.create_string_buffer(64)
            err = None
            try:
                cctx.check(cctx.L.mct_p2p_create(cctx.h, world, rank, hbuf))
            except Exception as e:                      # noqa: BLE001
                err = "create: %r" % (e,)
            if all_ok(err is None):
                allh = [None] * world
                dist.all_gather_object(allh, bytes(hbuf.raw))
                try:
                    cctx.check(cctx.L.mct_p2p_open(cctx.h, b"".join(allh)))
                    cam._p2p_world = world
                except Exception as e:                  # noqa: BLE001
                    err = "open: %r" % (e,)
                if all_ok(err is None):
                    barrier()
                    try:
                        for _ in range(5):
                            p2p_pts = cam.exchange_foot_points()
                        tw = time.perf_counter()
                        for _ in range(50):
                            p2p_pts = cam.exchange_foot_points()
                        tw = time.perf_counter() - tw
                    except Exception as e:              # noqa: BLE001
                        err = "exchange: %r" % (e,)
                    if all_ok(err is None):
                        fuser["p2p_us_per_frame"] = allmax(1e6 * tw / 50)
                        fuser["p2p_verified"] = bool(np.array_equal(p2p_pts, allp.cpu().numpy()))
                        fuser["p2p_what"] = ("k_foot_points_push (NVLink peer stores into all cameras' buffers) + k_foot_points_wait, "
                                             "host-timed incl. the read-out")
            if "p2p_us_per_frame" not in fuser:
                fuser["p2p_skipped"] = err or "a peer rank failed"
        except Exception as e:                          # noqa: BLE001
            fuser["p2p_skipped"] = "unexpected: %r" % (e,)

    ms_max, ms_e2e_max, ms_full_max = allmax(ms), allmax(ms_e2e), allmax(ms_full)
    scored_all, scored_e2e_all = allsum(scored), allsum(scored_e2e)
    launches_all = allsum(launches)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # dominant kernel by total event time over the profiled steps
        tot = sum(v[0] for v in prof.values()) or 1.0
        dom = max(prof.items(), key=lambda kv: kv[1][0])
        dom_name, (dom_ms, dom_n) = dom
        avg_us = 1e3 * dom_ms / max(dom_n, 1)
        rows, cols = seq.h0, seq.w0
        # ALGORITHMIC bytes per launch: what the reference's loads / stores amount to
        n_tot = N_OBJ * N_PART
        k_sel = CONFIG["selected"]
        k_haar = int(np.mean([(cam.selectors(o) < N_HAAR).sum() for o in range(N_OBJ)]))      # selected Haar features (measured)
        alg_all = {
            # B_mdch = N*rows*cols*3 + N*nb^3*4, plus the predict / test-set prologue fused into the kernel (68 B per particle)
            "k_mdch_sel": n_tot * (rows * cols * 3 + (N_BINS ** 3) * 4 + 68),
            "k_classify_staged": n_tot * (k_haar * 4 * 16 + (k_sel - k_haar) * 4 + 4),        # E[R_f] = 4 rects x 4 corners x 4 B
            "k_frame_prep": W * H + (W + 1) * (H + 1) * 4 + 3 * W * H + 2 * W * H,            # integral + joint-bin image
            "k_pf_update": n_tot * 44,
            "k_pf_predict_testset": n_tot * 68,
        }
        alg = alg_all.get(dom_name, 0)
        traffic, traffic_all = None, {}
        try:
            traffic_all = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            traffic = traffic_all.get(dom_name)
        except Exception:
            pass
        achieved = alg / (avg_us * 1e-6) / 1e9 if avg_us > 0 else 0.0
        # on-chip ceilings: the gather kernels work out of shared memory / L2, so their traffic there (ncu counters of
        # the committed capture, per launch) is set against this pool's measured shared-memory and L2 read peaks (tools/peaks_lab.cu)
        onchip = None
        try:
            pk = json.load(open(os.path.join(ROOT, "profiles", "r01_peaks.json")))
            oc = json.load(open(os.path.join(ROOT, "profiles", "roofline_onchip.json"))).get(dom_name)
            if oc and avg_us > 0:
                smem_b = oc["smem_wavefronts"] * 128.0
                onchip = {"smem_peak_gbs": pk["smem_read_gbs"], "l2_peak_gbs": pk["l2_read_gbs"], "peaks_source": "profiles/r01_peaks.json (tools/peaks_lab.cu)",
                          "smem_bytes_per_launch": smem_b, "smem_achieved_gbs": smem_b / (avg_us * 1e-6) / 1e9,
                          "smem_frac": smem_b / (avg_us * 1e-6) / 1e9 / pk["smem_read_gbs"],
                          "smem_bank_conflict_share": oc["smem_bank_conflict_wavefronts"] / max(oc["smem_wavefronts"], 1.0),
                          "l2_read_bytes_per_launch": oc.get("l2_bytes"),
                          "l2_achieved_gbs": (oc["l2_bytes"] / (avg_us * 1e-6) / 1e9) if oc.get("l2_bytes") else None,
                          "counters_source": "profiles/roofline_onchip.json (ncu --set full, per launch)"}
        except Exception:
            pass
        shares = {k: round(v[0] / tot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        # what really bounds the dominant kernel (ncu: DRAM traffic is ~1 % of the algorithmic bytes, everything is staged on chip)
        bound = {"k_mdch_sel": "smem", "k_classify_staged": "smem/latency", "k_pf_update": "latency", "k_frame_prep": "hbm",
                 "k_haar_matrix_batch": "l1/l2", "k_pf_predict_testset": "latency"}.get(dom_name, "hbm")
        out = {
            "metric": METRIC, "value": scored_all / (ms_max * 1e-3), "unit": "particles/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "e2e": {"value": scored_e2e_all / (ms_e2e_max * 1e-3), "unit": "particles/s",
                    "h2d_bytes_per_step": (1 if gray_only else (4 if (not args.e2e_bgr) else 3)) * W * H + N_OBJ * 4 * N_PART * 4,
                    "input": ("gray u8 plane (Haar-only classifiers read nothing else)" if gray_only else "4 u8 planes (gray,R,G,B)") if (not args.e2e_bgr) else "packed BGR frame (split + gray conversion on the device)", "d2h_bytes_per_step": N_OBJ * C.sizeof(capi.PfSummary),
                    "ms_per_step": ms_e2e_max / args.steps,
                    "call": ("mcth_camera_step, streaming (phases 1 | 16): frame t's boxes are returned after frame t + 1 has been queued; "
                             "one frame uploaded, tracked and read back per call") if stream_e2e else "mcth_camera_step, synchronous per frame"},
            "tracked_frames_per_sec_per_camera": args.steps / (ms_full_max * 1e-3),
            "full_frame": {"ms_per_frame": ms_full_max / args.steps,
                           "training_sets": ("device: k_train_gen_default behind the frame's k_pf_update, tracker streams device-resident, no host round trip"
                                             if mhost.device_training_sets() else "host enumeration (MCT_HOST_TRAINING_SETS=1)"),
                           "kernel_us_per_frame": {k: round(1e3 * v[0] / n_pf, 2) for k, v in sorted(prof_full.items(), key=lambda kv: -kv[1][0])}},
            "gpu_launches": int(launches_all),
            "roofline": {"kernel": dom_name, "bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "achieved_is": "ALGORITHMIC bytes per launch / launch time (SURVEY 8d), set against the HBM copy peak as the contract asks; "
                                        "the limiter named in `bound` and its own fraction are in `onchip`",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                         "avg_launch_us": avg_us, "algorithmic_bytes_per_launch": alg, "kernel_time_shares": shares, "onchip": onchip,
                         "kernels": {k: {"avg_launch_us": round(1e3 * v[0] / max(v[1], 1), 2), "algorithmic_bytes": alg_all.get(k, 0),
                                         "achieved_gbs": round(alg_all.get(k, 0) / max(1e3 * v[0] / max(v[1], 1), 1e-9) / 1e3, 1),
                                         "dram_traffic_bytes": traffic_all.get(k)} for k, v in prof.items()},
                         "note": "working set (~3 MB per camera) is L2/shared-memory resident: dram traffic << algorithmic bytes; "
                                 "k_pf_update is bound by three sequential f32 add chains (bit-exact resampling), not by bandwidth"},
            "clocks": sampler.summary(),
        }
        if fuser is not None:
            out["fuser_exchange"] = fuser
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_leg()
            out["cpu_baseline_full_frame"] = cpu_baseline_leg(phases=3)
        print(json.dumps(out))
    cam.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ configs[2] / [3]: the camera network
NET_POOL = 24          # frames per camera in the host pool of the network runs (uploaded every step: these runs are end to end)


def network_scene(n_cameras):
    from multicameratracking_b200 import synth
    c = CONFIGS[2] if CONFIG["workload"].startswith("configs[2]") else CONFIGS[3]
    return c, synth.MultiViewScene(n_cameras, N_OBJ, W, H, n_frames=NET_POOL, camouflage=c["camouflage"])


def run_network(args):
    """One step = one frame of the whole per-frame schedule of CameraNetwork::TrackObjectsOnCurrentFrame (CameraNetwork.cpp:537-613)
    for this rank's camera: frame upload, score path, fuser exchange INSIDE the step, feedback / re-weighting, UpdateClassifier,
    (appearance) global model refresh, state read-out.  Under torchrun: one camera per rank, exchange over NCCL; alone: all cameras
    of the configuration on this GPU (exchange in process)."""
    import torch
    import torch.distributed as dist
    from multicameratracking_b200 import synth
    from multicameratracking_b200 import host as mhost
    from multicameratracking_b200.network import TorchExchange

    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[args.config]
    n_cameras = world if world > 1 else cfg["cameras"]
    _, scene = network_scene(n_cameras)
    my = [rank] if world > 1 else list(range(n_cameras))
    total = args.steps + args.warmup
    SEG = 3 if cfg["net"].get("geometric_fusion") else 12      # frames per tracking segment, see build()
    frames = {c: [scene.frame(c, t) for t in range(SEG + 1)] for c in my}
    over = dict(feature_type=FEAT, n_haar=max(N_HAAR, 1), n_bins=N_BINS, weak_type=WEAK, strong_type=STRONG, percent_selected=PCT, n_particles=N_PART,
                sd_x=SD[0], sd_y=SD[1], sd_sx=SD[2], sd_sy=SD[3])
    ex = TorchExchange() if world > 1 else None
    state = {"cams": [], "net": None}

    def build():
        """(Re)initialise the trackers on frame 0 of the scene -- untimed.  The reference's geometric feedback is not a stable
        controller on long synthetic sequences (its rearrangement triggers whenever the average particle's foot, computed with
        h0 instead of h0 / 2, is 20 px from the fused estimate, ParticleFilterTracker.cpp:1093-1105, and rescales the particles
        by the offset: with the 40 x 80 objects of the synthetic specification it fires EVERY frame); the reference itself runs into
        its sampler / empty-training-set assertions after 4 (2 cameras) to 17 (4 cameras) frames of this scene, and so does this
        path, on the same frame.  The bench therefore tracks short segments from a fresh initialisation; every timed step is a
        regular frame of the schedule (with the appearance fuser, whose re-weighting starts at frame 3, segments are 12 frames)."""
        if state["net"] is not None:
            state["net"].close()
            for cam in state["cams"]:
                cam.close()
        cams = []
        for c in my:
            cam = mhost.Camera(local, n_frames=SEG + 4)
            cam.set_frame(*frames[c][0])
            for o in range(N_OBJ):
                cam.add_object(scene.box(c, 0, o), rng_seed=1000 * c + o + 1, **over)
            cams.append(cam)
        net = mhost.Network(cams, scene.homographies, camera_ids=my, n_cameras=n_cameras, exchange=ex.as_tuple() if ex else None, **cfg["net"])
        net.init()
        if ex:
            ex.bind_rows(net)
        state["cams"], state["net"] = cams, net

    noise = {c: [np.stack([synth.noise(N_PART, SD, camera=c, obj=o, t=k + 1) for o in range(N_OBJ)]) for k in range(SEG)] for c in my}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stages = {}
    timed = {"s": 0.0, "launches": 0}

    def step(i, record, after_build=None):
        k = i % SEG
        if k == 0:
            build()
            barrier()
            if after_build:
                after_build()
        cams, net = state["cams"], state["net"]
        l0 = sum(cam.launches for cam in cams)
        t0 = time.perf_counter()
        for j, c in enumerate(my):
            cams[j].set_frame(*frames[c][k + 1])
        net.track(k + 1, [noise[c][k] for c in my])
        for cam in cams:
            cam.sync()
        dt = time.perf_counter() - t0
        if record:
            timed["s"] += dt
            timed["launches"] += sum(cam.launches for cam in cams) - l0
            for kk, v in net.stage_ms().items():
                stages[kk] = stages.get(kk, 0.0) + v

    sampler = ClockSampler(local); sampler.start()
    for i in range(total):
        step(i, i >= args.warmup)
    barrier()
    sampler.stop_flag = True
    # per-kernel CUDA-event times of one more (untimed) segment of the first local camera: what the frame's device time is made of
    kernel_us = None
    if os.environ.get("MCT_BENCH_KERNELS", "1") == "1":
        from multicameratracking_b200 import capi
        Lc = capi.load()
        prof_acc, prof_frames = {}, 0
        holder = {}

        def profile_on():          # (the segment's first step re-creates the cameras: take the context of the NEW first camera)
            holder["ctx"] = C.c_void_p(state["cams"][0].L.mcth_camera_ctx(state["cams"][0].h))
            Lc.mct_profile_enable(holder["ctx"], 1)

        for i in range(SEG):
            step(i, False, after_build=profile_on)
            for kname, (ms_k, n_k) in capi.profile_read(Lc, holder["ctx"]).items():
                acc = prof_acc.setdefault(kname, [0.0, 0]); acc[0] += ms_k; acc[1] += n_k
            prof_frames += 1
        Lc.mct_profile_enable(holder["ctx"], 0)
        kernel_us = {k: [round(1e3 * v[0] / prof_frames, 1), round(v[1] / prof_frames, 1)] for k, v in sorted(prof_acc.items(), key=lambda kv: -kv[1][0])}
        barrier()
    dt = timed["s"]
    launches = timed["launches"]
    cams, net = state["cams"], state["net"]

    def allmax(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); return float(tt.item())

    def allsum(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.SUM); return float(tt.item())

    dt_max = allmax(dt); launches_all = allsum(launches)
    st = {k: allmax(v / args.steps) for k, v in sorted(stages.items())}
    if rank == 0:
        cam_frames = n_cameras * args.steps
        h2d = len(my) * (4 * W * H + N_OBJ * 4 * N_PART * 4)
        val = cam_frames / dt_max
        out = {"metric": "camera_frames_tracked_per_sec", "value": val, "unit": "camera-frames/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": 1e3 * dt_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": dict(CONFIG, cameras=n_cameras, cameras_per_gpu=len(my), fusers=cfg["net"],
                                                                 l2_policy="inputs uploaded from host memory every step (end to end)"),
               "timing": "host wall clock around every timed step with the camera streams drained at its end, summed over the steps, max over "
                         "ranks: the network's schedule synchronises with the device several times per frame (training-set generation, fuser "
                         "decisions), so the host clock is the device timeline; trackers are re-initialised (untimed) every %d frames" % SEG,
               "tracked_frames_per_sec_per_camera": args.steps / dt_max / (len(my) if world == 1 else 1),
               "particles_scored_per_sec": cam_frames * N_OBJ * N_PART / dt_max,
               "e2e": {"value": val, "unit": "camera-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": len(my) * N_OBJ * 80,
                       "note": "the network path IS end to end: host frames and noise in, boxes out, every step"},
               "stage_ms_per_frame": {k: round(v, 4) for k, v in st.items()},
               "kernel_us_per_frame_camera0": kernel_us,      # {kernel: [us per frame, launches per frame]}, CUDA events, rank 0's first camera
               "gpu_launches": int(launches_all), "roofline": None, "clocks": sampler.summary()}
        print(json.dumps(out))
    net.close()
    for cam in cams:
        cam.close()
    if world > 1:
        dist.destroy_process_group()


def run_network_reference(args):
    """The reference's own code (oracle/_ref: /root/reference/src compiled against header shims) running the same multi-camera schedule on
    the host CPU, single thread as the reference ships; a bounded sample of the workload (fewer cameras / objects when a frame takes long)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from multicameratracking_b200 import synth
    from oracle.ref_py import Ref, RefNetwork
    if not Ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built and /root/reference is absent"}))
        return
    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_cameras = world if world > 1 else cfg["cameras"]
    # bounded sample: configs[3] needs ~80 s per frame at its full shape
    sc, so = (n_cameras, N_OBJ) if args.config == 2 else (2, 4)
    steps, warm = (min(args.steps, 4), 1) if args.config == 2 else (2, 1)
    scene = synth.MultiViewScene(sc, so, W, H, n_frames=NET_POOL, camouflage=cfg["camouflage"])
    init = np.array([[scene.box(c, 0, o) for o in range(so)] for c in range(sc)], np.float32)
    seeds = np.array([[1000 * c + o + 1 for o in range(so)] for c in range(sc)])
    ref = Ref()
    net = RefNetwork(ref, sc, so, steps + warm + 2, scene.homographies, init, seeds, feature_type=FEAT, n_haar=max(N_HAAR, 1), strong_type=STRONG,
                     weak_type=WEAK, n_particles=N_PART, **cfg["net"])
    net.init([scene.frame(c, 0) for c in range(sc)])
    nz = lambda t: np.stack([np.stack([synth.noise(N_PART, SD, camera=c, obj=o, t=t) for o in range(so)]) for c in range(sc)])   # noqa: E731
    for i in range(warm):
        net.track(i + 1, [scene.frame(c, pingpong(i, NET_POOL)) for c in range(sc)], nz(i + 1))
    t0 = time.perf_counter()
    for i in range(warm, warm + steps):
        net.track(i + 1, [scene.frame(c, pingpong(i, NET_POOL)) for c in range(sc)], nz(i + 1))
    dt = time.perf_counter() - t0
    net.close()
    # camera-frames/s normalised to the configuration's objects per camera
    val = sc * steps * (so / float(N_OBJ)) / dt
    sample = "%d frames x %d cameras x %d of %d objects x %d particles, full schedule with fusers, reference build single thread" % (steps, sc, so, N_OBJ, N_PART)
    print(json.dumps({"impl": "reference", "metric": "camera_frames_tracked_per_sec", "value": val, "unit": "camera-frames/s", "n_gpus": args.gpus,
                      "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": dict(CONFIG, cameras=n_cameras),
                      "cpu_baseline": {"value": val, "unit": "camera-frames/s", "cores": 1, "kind": "reference", "sample": sample},
                      "e2e": {"value": val, "unit": "camera-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ configs[4]: Haar scoring sweep
def run_sweep(args):
    """roofline characterisation of the Haar rectangle-sum gather (FeatureVector::Compute): N boxes x F features x frame size"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.impl == "reference":
        print(json.dumps({"impl": "reference", "unavailable": "configs[4] is a roofline sweep of the GPU kernel; the CPU arm of the path is --config 0 / 1"}))
        return
    import torch
    from multicameratracking_b200 import capi, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    ctx = capi.Context(0)
    L = capi.load()
    peak = 6545.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    pts = []
    for (w, h) in [(640, 480), (1280, 720), (1920, 1080)]:
        seq = synth.Sequence(w, h, 1)
        ctx.frame_set(seq.frame(0)[0])
        for F in (250, 2000):
            arrs = synth.haar_table(1, F, seq.w0, seq.h0)
            sumR = int(arrs[0].sum())
            clf = capi.StrongClassifier(ctx, capi.FEAT_HAAR, seq.w0, seq.h0, n_haar=F, n_selected=max(1, F // 5), haar=arrs)
            for N in (1000, 8000, 64000):
                rng = np.random.default_rng(5)
                bx = capi.make_boxes(rng.integers(1, w - seq.w0 - 2, N).astype(np.int32), rng.integers(1, h - seq.h0 - 2, N).astype(np.int32),
                                     np.ones(N, np.float32), np.ones(N, np.float32))
                for _ in range(max(3, args.warmup)):
                    clf.features(bx)
                capi.profile_enable(L, ctx.h, True)
                for _ in range(5):
                    clf.features(bx)
                prof = capi.profile_read(L, ctx.h)
                capi.profile_enable(L, ctx.h, False)
                ms = sum(v[0] for k, v in prof.items() if k.startswith("k_haar") or k.startswith("k_tile"))
                n = max(v[1] for k, v in prof.items() if k.startswith("k_haar"))
                us = 1e3 * ms / n
                alg = N * sumR * 16 + N * F * 4
                pts.append({"frame": [w, h], "N": N, "F": F, "us": round(us, 2), "kernels": sorted(k for k in prof if k.startswith("k_haar") or k.startswith("k_tile")),
                            "algorithmic_bytes": alg, "achieved_gbs": round(alg / us / 1e3, 1), "frac_of_hbm_peak": round(alg / us / 1e3 / peak, 3),
                            "rect_sums_per_s_G": round(N * sumR / us / 1e3, 2)})
            clf.close()
    best = max(pts, key=lambda q: q["achieved_gbs"])
    print(json.dumps({"metric": "haar_rect_sums_per_sec", "value": best["rect_sums_per_s_G"] * 1e9, "unit": "rect-sums/s", "n_gpus": 1, "steps": 5,
                      "warmup": max(3, args.warmup), "ms_per_step": best["us"] * 1e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "configs[4]: Haar scoring sweep, N boxes x F features x frame size, uniformly random boxes", "best_point": best},
                      "roofline": {"kernel": "k_haar_matrix(_tiled)", "bound": "smem/L2 (staged gather) - reported against the HBM copy peak in algorithmic bytes",
                                   "achieved": best["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": best["frac_of_hbm_peak"], "traffic": None},
                      "points": pts, "gpu_launches": 5 * len(pts)}))
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-bgr", action="store_true", help="e2e leg uploads the packed BGR frame (split + gray on the device) instead of 4 planes")
    ap.add_argument("--config", type=int, default=1, choices=[0, 1, 2, 3, 4],
                    help="BASELINE.json configs[k]; 1 (default) is the configuration the headline metric is quoted on")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.config == 4:
        return run_sweep(args)
    configure(args.config)
    if args.config in (2, 3):
        return run_network_reference(args) if args.impl == "reference" else run_network(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
