"""Cross-camera exchange for the fusers: one camera per rank per GPU.

The reference's CameraNetwork reads the other cameras' state through pointers inside one process
(CameraNetwork.cpp:182-216 foot points for the geometric fuser, :276-320 training SampleSets for the appearance
fuser).  With cameras sharded across GPUs the same payloads travel with ONE all-gather per frame:

  geometric fuser   foot point of the highest-weight particle per object, float32 [n_obj][2], packed on the device by
                    mct_fuser_pack_foot_points (ParticleFilterTracker.cpp:1223-1238); every rank then runs the tiny
                    per-object fusion redundantly, so no broadcast is needed.
  appearance fuser  per-object training FEATURE ROWS (instead of image pointers): float32 [F][P] and [F][Nn], padded to
                    the largest count over ranks, plus the counts; each rank pools them in camera order and feeds
                    mct_classifier_update_from_features (AppearanceBasedInformationFuser.cpp:44-90).

Works on CUDA tensors with the NCCL backend (NVLink / NVSwitch) and on CPU tensors with gloo (tests).  There is no
compute here: only packing, the collective and unpacking.
"""
import torch
import torch.distributed as dist


class CameraNetworkExchange:
    def __init__(self, n_objects, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (one process per camera / GPU)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n_objects = int(n_objects)
        self.device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"      # where the collectives' tensors live

    # ------------------------------------------------------------------ geometric fuser payload
    def allgather_foot_points(self, local):
        """local: float32 [n_obj][2] (this camera).  Returns [world][n_obj][2], index 0 = camera id = rank."""
        local = local.contiguous()
        if tuple(local.shape) != (self.n_objects, 2) or local.dtype != torch.float32:
            raise ValueError("foot points must be float32 [%d][2]" % self.n_objects)
        return self._allgather(local)

    def _allgather(self, local):
        """all_gather_into_tensor with the output concatenated along dim 0 (the form both NCCL and gloo accept)"""
        local = local.contiguous()
        flat = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(flat, local, group=self.group)
        return flat.view((self.world,) + tuple(local.shape))

    # ------------------------------------------------------------------ appearance fuser payload
    def allgather_feature_rows(self, fpos, fneg):
        """fpos: float32 [F][P], fneg: float32 [F][Nn] of ONE object on this camera (P, Nn may differ per rank).
        Returns (pooled_pos [F][sum P], pooled_neg [F][sum Nn], counts int64 [world][2]) in camera order."""
        F = fpos.shape[0]
        if fneg.shape[0] != F:
            raise ValueError("positive and negative rows must have the same feature count")
        dev = fpos.device
        counts = torch.tensor([fpos.shape[1], fneg.shape[1]], dtype=torch.int64, device=dev)
        allc = self._allgather(counts.view(1, 2)).view(self.world, 2)
        pmax, nmax = int(allc[:, 0].max()), int(allc[:, 1].max())
        pad = torch.zeros((F, pmax + nmax), dtype=torch.float32, device=dev)
        pad[:, :fpos.shape[1]] = fpos
        pad[:, pmax:pmax + fneg.shape[1]] = fneg
        allp = self._allgather(pad)
        pos = torch.cat([allp[r, :, :int(allc[r, 0])] for r in range(self.world)], dim=1)
        neg = torch.cat([allp[r, :, pmax:pmax + int(allc[r, 1])] for r in range(self.world)], dim=1)
        return pos.contiguous(), neg.contiguous(), allc

    # ------------------------------------------------------------------ timing helper (bench.py)
    def max_over_ranks(self, value, device=None):
        t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def sum_over_ranks(self, value, device=None):
        t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())


class GeometricFusion:
    """CameraNetwork::FuseGeometrically + FeedbackInformationFromGeometricFusion (CameraNetwork.cpp:94-268) with one camera
    per rank: pack this camera's foot points on the device, ONE all-gather, then every rank runs the per-object fuser
    (principal-axis intersection + Kalman; tiny, host side) on identical inputs and feeds the result back into its own
    camera's particles (device: ground pdf, rearrangement, re-score)."""

    def __init__(self, exchange, camera, homographies, p2p=False):
        """p2p: exchange the foot points with the library's own peer-memory kernels (NVLink stores into every camera's
        receive buffer, fused with the foot-point kernel) instead of an NCCL all-gather"""
        from .host import GeometricFuser
        self.p2p = bool(p2p)
        if self.p2p:
            def gather(b):
                out = [None] * exchange.world
                dist.all_gather_object(out, b, group=exchange.group)
                return out
            camera.p2p_setup(exchange.rank, exchange.world, gather)
            dist.barrier(group=exchange.group)
        self.ex = exchange
        self.cam = camera
        self.H = [torch.as_tensor(h, dtype=torch.float64).reshape(3, 3).numpy() for h in homographies]
        if len(self.H) != exchange.world:
            raise ValueError("one homography per camera (rank) is needed")
        self.fusers = [GeometricFuser(self.H) for _ in range(exchange.n_objects)]
        self._foot = None

    def step(self, device=None):
        """one frame: returns (foot points [world][n_obj][2], [(rearranged, distance)] per object of this camera).
        self.last_ms holds the host wall time of the four stages (pack + all-gather incl. the read-back, fusers, feedback)."""
        import time
        from .host import fuse_many
        n = self.ex.n_objects
        if self._foot is None:
            self._foot = torch.zeros((n, 2), dtype=torch.float32, device=device or "cuda")
        t0 = time.perf_counter()
        if self.p2p:
            allp = self.cam.exchange_foot_points()
        else:
            self.cam.pack_foot_points(self._foot.data_ptr())
            allp = self.ex.allgather_foot_points(self._foot).cpu().numpy()
        t1 = time.perf_counter()
        means, covs, fresh = fuse_many(self.fusers, allp)
        t2 = time.perf_counter()
        if fresh.all():
            out = [(False, -1.0)] * n              # the frame that initialises the filters gives no feedback yet
        elif hasattr(self.cam, "ground_feedback_all") and not fresh.any():
            out = self.cam.ground_feedback_all(self.H[self.ex.rank], means, covs)
        else:
            out = [(False, -1.0) if fresh[o] else self.cam.ground_feedback(o, self.H[self.ex.rank], means[o], covs[o]) for o in range(n)]
        t3 = time.perf_counter()
        self.last_ms = {"exchange": 1e3 * (t1 - t0), "fusers": 1e3 * (t2 - t1), "feedback": 1e3 * (t3 - t2)}
        return allp, out

    def close(self):
        for f in self.fusers:
            f.close()


class AppearanceFusion:
    """AppearanceBasedInformationFuser of ONE object on ONE camera (rank) with the other cameras' training samples arriving
    as feature rows (AppearanceBasedInformationFuser.cpp:17-121, CameraNetwork.cpp:276-320).

      global_clf   this rank's global appearance classifier of the object (capi.StrongClassifier made with
                   learning_rate = 0, :28).  Every rank must build it from the SAME feature table (same Haar draws): rows
                   computed on one camera are only meaningful to a classifier with the same features.
      learn()      LearnGlobalAppearanceModel: feature rows of the local positive / negative boxes under the global
                   classifier's features -> ONE all-gather -> pooled in camera order -> each sample kept iff
                   randfloat() > 0.5 (positives first, then negatives) -> Update from the pooled rows
      fuse()       FuseInformation + Object::UpdateParticleWeightUsingMultiCameraAppearanceModel (Object.cpp:496-528) +
                   ForceParticleResampling (Object.cpp:445): particles scored in place by the global classifier, weight *=
                   sigmoid(H), ResampleIfRequired, then ResampleParticles(false)
    """

    def __init__(self, exchange, ctx, global_clf):
        self.ex = exchange
        self.ctx = ctx
        self.clf = global_clf
        self.trained = False

    def learn(self, pos_boxes, neg_boxes, rng_next_float):
        """pos_boxes / neg_boxes: capi.make_boxes(...) of this camera's training sets for the object (current frame set).
        With NCCL the rows never leave the device: feature kernels -> all-gather -> column selection -> update."""
        dev = self.ex.device
        on_device = dev == "cuda" and hasattr(self.clf, "features_dev") and hasattr(pos_boxes, "n")
        if on_device:
            F, P, Nn = self.clf.F, int(pos_boxes.n), int(neg_boxes.n)
            fpos = torch.empty((F, P), dtype=torch.float32, device="cuda")
            fneg = torch.empty((F, Nn), dtype=torch.float32, device="cuda")
            torch.cuda.current_stream().synchronize()          # the C-ABI works on the camera's stream
            self.clf.features_dev(pos_boxes, fpos.data_ptr(), P)
            self.clf.features_dev(neg_boxes, fneg.data_ptr(), Nn)
        else:
            fpos = torch.from_numpy(self.clf.features(pos_boxes)).to(dev)
            fneg = torch.from_numpy(self.clf.features(neg_boxes)).to(dev)
        pos, neg, counts = self.ex.allgather_feature_rows(fpos, fneg)
        kp = drop_samples(rng_next_float, pos.shape[1])
        kn = drop_samples(rng_next_float, neg.shape[1])
        if not kp or not kn:
            raise RuntimeError("appearance fusion: every pooled %s sample was dropped" % ("positive" if not kp else "negative"))
        if on_device:
            sp = pos[:, torch.as_tensor(kp, device="cuda")].contiguous()
            sn = neg[:, torch.as_tensor(kn, device="cuda")].contiguous()
            torch.cuda.current_stream().synchronize()
            self.clf.update_from_features_dev(sp.data_ptr(), len(kp), len(kp), sn.data_ptr(), len(kn), len(kn))
        else:
            self.clf.update_from_features(pos[:, kp].contiguous().cpu().numpy(), neg[:, kn].contiguous().cpu().numpy())
        self.trained = True
        return counts, kp, kn

    def fuse(self, pf, w0, h0, u01_update, u01_forced):
        """returns the mct_pf_summary of the re-weighting call"""
        from . import capi
        if not self.trained:
            raise RuntimeError("appearance fusion: learn() has not run")
        summ = capi.track_score(self.ctx, [pf], [self.clf], [(w0, h0, 0, 0, 1)], None, [u01_update])[0]
        pf.ResampleParticles(False, u01_forced)
        return summ


def drop_samples(rng_next_float, n, threshold=0.5):
    """AppearanceBasedInformationFuser.cpp:66-82: a pooled sample is kept iff randfloat() > RANDOM_SAMPLE_SELECTION_THRESHOLD
    (0.5, :7), in order (all positives, then all negatives), one draw per sample from the MWC stream
    (rng_next_float() -> float in [0,1))."""
    return [i for i in range(n) if rng_next_float() > threshold]


class TorchExchange:
    """NetworkExchange for host.Network over torch.distributed: rank r holds camera r (one process per camera per GPU).

    Small host payloads (foot points, counts) travel as CUDA (NCCL) or CPU (gloo) tensors; the appearance fuser's feature rows
    are all-gathered between two device buffers this object owns (torch tensors, so that NCCL can use them) and hands to the
    network with bind_rows()."""

    def __init__(self, group=None, device=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (one process per camera / GPU)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.cuda = dist.get_backend(group) == "nccl"
        self.device = device if device is not None else ("cuda" if self.cuda else "cpu")
        self._small = {}
        self.send = self.recv = None

    def as_tuple(self):
        return (self.rank, self.world, self.allgather_host, self.alltoall_rows)

    def allgather_host(self, send_ptr, recv_ptr, nbytes):
        import ctypes as C
        nbytes = int(nbytes)
        if nbytes not in self._small:
            pin = self.cuda
            self._small[nbytes] = (torch.empty(nbytes, dtype=torch.uint8, pin_memory=pin), torch.empty(nbytes * self.world, dtype=torch.uint8, pin_memory=pin),
                                   torch.empty(nbytes, dtype=torch.uint8, device=self.device), torch.empty(nbytes * self.world, dtype=torch.uint8, device=self.device))
        hs, hr, ds, dr = self._small[nbytes]
        C.memmove(hs.data_ptr(), send_ptr, nbytes)
        if self.cuda:
            ds.copy_(hs, non_blocking=True)
            dist.all_gather_into_tensor(dr, ds, group=self.group)
            hr.copy_(dr, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            dist.all_gather_into_tensor(hr, hs, group=self.group)
        C.memmove(recv_ptr, hr.data_ptr(), nbytes * self.world)

    def bind_rows(self, network):
        """allocate the feature-row exchange buffers and give them to the network (after network.init())"""
        n = network.row_bytes()
        if n == 0:
            return
        # both buffers hold one slot per camera: send slot L = this camera's rows of learner round L, recv slot c = camera c's rows
        # of the round in which this camera learns
        self.send = torch.empty(n, dtype=torch.uint8, device=self.device)
        self.recv = torch.empty(n, dtype=torch.uint8, device=self.device)
        network.set_row_buffers(self.send.data_ptr(), self.recv.data_ptr())

    def alltoall_rows(self, slot_bytes, stream_ptr):
        """recv slot c <- camera c's send slot `rank`: one collective per frame for the appearance fuser's feature rows"""
        if self.cuda:
            st = torch.cuda.ExternalStream(int(stream_ptr))
            with torch.cuda.stream(st):
                dist.all_to_all_single(self.recv, self.send, group=self.group)          # ordered on the camera's own stream
        else:
            # gloo has no all-to-all: gather everything and keep the slots addressed to this rank (CPU tests, small worlds)
            slot = int(slot_bytes)
            full = torch.empty(self.send.numel() * self.world, dtype=torch.uint8)
            dist.all_gather_into_tensor(full, self.send, group=self.group)
            per = self.send.numel()
            for c in range(self.world):
                self.recv[c * slot:(c + 1) * slot] = full[c * per + self.rank * slot: c * per + (self.rank + 1) * slot]
