"""Multi-camera host logic on CPU: world_size-2 gloo groups (one process per camera), 127.0.0.1 rendezvous.
Covers the cross-camera exchange of multicameratracking_b200/network.py and the bench's max-over-ranks
reduction; no GPU and no CUDA call."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multicameratracking_b200.network import CameraNetworkExchange, drop_samples
        n_obj, F = 3, 7
        ex = CameraNetworkExchange(n_obj)
        assert (ex.rank, ex.world) == (rank, world)
        # geometric fuser: foot points, camera id = rank
        local = torch.arange(n_obj * 2, dtype=torch.float32).reshape(n_obj, 2) + 100.0 * rank
        allp = ex.allgather_foot_points(local)
        assert tuple(allp.shape) == (world, n_obj, 2)
        for r in range(world):
            assert torch.equal(allp[r], torch.arange(n_obj * 2, dtype=torch.float32).reshape(n_obj, 2) + 100.0 * r)
        with pytest.raises(ValueError):
            ex.allgather_foot_points(torch.zeros(n_obj + 1, 2))
        # appearance fuser: ragged feature rows, pooled in camera order
        g = torch.Generator().manual_seed(1234)
        P = [5, 2][rank % 2] + rank
        Nn = [1, 4][rank % 2]
        rows = {r: (torch.rand(F, [5, 2][r % 2] + r, generator=torch.Generator().manual_seed(10 + r)),
                    torch.rand(F, [1, 4][r % 2], generator=torch.Generator().manual_seed(20 + r))) for r in range(world)}
        fpos, fneg = rows[rank]
        assert fpos.shape[1] == P and fneg.shape[1] == Nn
        pos, neg, counts = ex.allgather_feature_rows(fpos, fneg)
        assert torch.equal(pos, torch.cat([rows[r][0] for r in range(world)], dim=1))
        assert torch.equal(neg, torch.cat([rows[r][1] for r in range(world)], dim=1))
        assert counts.tolist() == [[rows[r][0].shape[1], rows[r][1].shape[1]] for r in range(world)]
        # timing reduction used by bench.py: max over ranks, sum over ranks
        assert ex.max_over_ranks(1.0 + rank) == float(world)
        assert ex.sum_over_ranks(10.0) == 10.0 * world
        # random sample dropping of the appearance fuser follows the MWC stream in order
        seq = iter([0.9, 0.1, 0.6, 0.5, 0.51])
        assert drop_samples(lambda: next(seq), 5) == [0, 2, 4]
        q.put((rank, "ok"))
    except BaseException as e:           # noqa: BLE001 - report to the parent
        q.put((rank, "FAIL %r" % (e,)))
        raise
    finally:
        dist.destroy_process_group()


class _StandInCamera:
    """duck-typed Camera for the CPU test: foot points written at the given address, feedback calls recorded"""

    def __init__(self, foot):
        self.foot = np.ascontiguousarray(foot, np.float32)
        self.calls = []

    def pack_foot_points(self, ptr):
        import ctypes
        ctypes.memmove(ptr, self.foot.ctypes.data, self.foot.nbytes)

    def ground_feedback(self, o, H, mean, cov):
        self.calls.append((o, np.array(H), np.array(mean), np.array(cov)))
        return False, 0.0


def _fusion_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multicameratracking_b200.network import CameraNetworkExchange, GeometricFusion
        from oracle import fuser_oracle as fo
        n_obj = 2
        H = [np.eye(3), np.array([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])]
        ex = CameraNetworkExchange(n_obj)
        feet = lambda r, t: np.array([[100.0 + 10 * r + 3 * t, 50.0], [200.0 - 5 * r + t, 80.0]], np.float32)   # noqa: E731
        cam = _StandInCamera(feet(rank, 0))
        gf = GeometricFusion(ex, cam, H)
        kfs = [None] * n_obj
        for t in range(4):
            cam.foot = feet(rank, t)
            allp, out = gf.step(device="cpu")
            assert allp.shape == (world, n_obj, 2)
            for r in range(world):
                np.testing.assert_array_equal(allp[r], feet(r, t))
            for o in range(n_obj):
                g = fo.principal_axis_intersection(allp[:, o, :], H)
                if kfs[o] is None:
                    kfs[o] = fo.KalmanCV(np.float32(g[0]), np.float32(g[1]))
                    assert out[o] == (False, -1.0)
                else:
                    kfs[o].update(np.array([g[0], g[1]], np.float32), np.array([[5, 0], [0, 5]], np.float32))
            if t > 0:                 # every rank fed its own camera with the SAME posterior, computed redundantly
                last = cam.calls[-n_obj:]
                for o, (oo, Hc, mean, cov) in enumerate(last):
                    assert oo == o and np.array_equal(Hc, H[rank])
                    np.testing.assert_allclose(mean, kfs[o].mean(), rtol=1e-5, atol=1e-4)
                    np.testing.assert_allclose(cov, kfs[o].cov(), rtol=1e-5, atol=1e-4)
        assert len(cam.calls) == 3 * n_obj
        gf.close()
        q.put((rank, "ok"))
    except BaseException as e:           # noqa: BLE001 - report to the parent
        q.put((rank, "FAIL %r" % (e,)))
        raise
    finally:
        dist.destroy_process_group()


class _StandInClassifier:
    """duck-typed capi.StrongClassifier: feature rows are a deterministic function of the box columns"""

    def __init__(self, F):
        self.F = F
        self.updated = None

    def features(self, boxes):
        cols = np.asarray(boxes, np.float32)
        return (np.arange(self.F, dtype=np.float32)[:, None] * 1000.0 + cols[None, :]).astype(np.float32)

    def update_from_features(self, fpos, fneg):
        self.updated = (np.array(fpos), np.array(fneg))


def _appearance_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multicameratracking_b200.network import AppearanceFusion, CameraNetworkExchange
        F = 5
        ex = CameraNetworkExchange(1)
        assert ex.device == "cpu"
        clf = _StandInClassifier(F)
        af = AppearanceFusion(ex, None, clf)
        cols = {0: ([1, 2, 3, 4], [50, 51]), 1: ([11, 12], [60, 61, 62])}     # ragged per camera
        draws = [0.9, 0.1, 0.6, 0.5, 0.51, 0.7, 0.2, 0.95, 0.3, 0.8, 0.99]   # 6 positives then 5 negatives
        it = iter(draws)
        counts, kp, kn = af.learn(cols[rank][0], cols[rank][1], lambda: next(it))
        assert counts.tolist() == [[4, 2], [2, 3]]
        assert kp == [0, 2, 4, 5] and kn == [1, 3, 4]
        allpos = np.array(cols[0][0] + cols[1][0], np.float32); allneg = np.array(cols[0][1] + cols[1][1], np.float32)
        np.testing.assert_array_equal(clf.updated[0][0], allpos[kp])        # feature row 0 carries the column itself
        np.testing.assert_array_equal(clf.updated[1][0], allneg[kn])
        assert clf.updated[0].shape == (F, 4) and clf.updated[1].shape == (F, 3)
        q.put((rank, "ok"))
    except BaseException as e:           # noqa: BLE001 - report to the parent
        q.put((rank, "FAIL %r" % (e,)))
        raise
    finally:
        dist.destroy_process_group()


def test_appearance_fusion_pooling_world2():
    """LearnGlobalAppearanceModel with one camera per rank (gloo, CPU): ragged per-camera training rows pooled in camera
    order on every rank, then thinned by the randfloat() > 0.5 rule, positives first"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_appearance_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_geometric_fusion_world2():
    """FuseGeometrically + feedback with one camera per rank (gloo, CPU): the all-gathered foot points drive identical
    per-object fusers on every rank; checked against the numpy oracle of the fuser arithmetic"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fusion_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_camera_network_exchange_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_bench_shards_cameras_by_rank():
    """one camera per rank: synthetic sequences and noise streams differ per camera and are reproducible"""
    from multicameratracking_b200 import synth
    a0 = synth.Sequence(320, 240, 2, camera=0, n_frames=4, obj_wh=(20, 30)).frame(1)[0]
    a1 = synth.Sequence(320, 240, 2, camera=1, n_frames=4, obj_wh=(20, 30)).frame(1)[0]
    b0 = synth.Sequence(320, 240, 2, camera=0, n_frames=4, obj_wh=(20, 30)).frame(1)[0]
    assert np.array_equal(a0, b0) and not np.array_equal(a0, a1)
    assert not np.array_equal(synth.noise(16, (1, 1, 1, 1), camera=0, obj=0, t=1), synth.noise(16, (1, 1, 1, 1), camera=1, obj=0, t=1))


class _StandInNetwork:
    """what TorchExchange.bind_rows needs from host.Network"""

    def __init__(self, nbytes):
        self.nbytes = nbytes
        self.ptrs = None

    def row_bytes(self):
        return self.nbytes

    def set_row_buffers(self, send, recv):
        self.ptrs = (send, recv)


def _torch_exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        from multicameratracking_b200.network import TorchExchange
        ex = TorchExchange()
        r, w, ag_host, a2a_rows = ex.as_tuple()
        assert (r, w) == (rank, world)
        # host payload through raw pointers, as the C++ CameraNetwork calls it (foot points: 8 objects x 2 floats; counts)
        for n in (64, 24):
            send = (np.arange(n, dtype=np.uint8) + 100 * rank).astype(np.uint8)
            recv = np.zeros(n * world, np.uint8)
            ag_host(send.ctypes.data, recv.ctypes.data, n)
            for k in range(world):
                assert np.array_equal(recv[k * n:(k + 1) * n], (np.arange(n, dtype=np.uint8) + 100 * k).astype(np.uint8))
        # feature rows: the exchange owns the buffers and hands their addresses to the network
        slot = 2048
        net = _StandInNetwork(slot * world)
        ex.bind_rows(net)
        assert net.ptrs == (ex.send.data_ptr(), ex.recv.data_ptr())
        # send slot L carries (rank, L); after the all-to-all recv slot c must carry (c, rank)
        rows = np.concatenate([np.full(slot, 16 * rank + L, np.uint8) for L in range(world)])
        C.memmove(net.ptrs[0], rows.ctypes.data, slot * world)
        a2a_rows(slot, 0)
        got = np.frombuffer((C.c_uint8 * (slot * world)).from_address(net.ptrs[1]), np.uint8)
        for c in range(world):
            assert (got[c * slot:(c + 1) * slot] == 16 * c + rank).all()
        q.put((rank, "ok"))
    except BaseException as e:           # noqa: BLE001
        q.put((rank, "FAIL %r" % (e,)))
        raise
    finally:
        dist.destroy_process_group()


def test_torch_exchange_two_ranks_gloo():
    """the NetworkExchange implementation bench.py and tools/ use under torchrun, on CPU tensors with gloo"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_torch_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
