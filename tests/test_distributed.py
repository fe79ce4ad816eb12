"""Landmark sharding: host-side logic with world_size 2 over gloo on CPU, and the NCCL data path on
>= 2 GPUs when the box has them."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(kind):
    from openslam_g2o_b200 import synth
    return synth.venice_like(30, 1500, seed=9) if kind == "ba" else synth.expmap_ba(30, 1500, seed=9)


def _cpu_worker(rank, world, port, out, kind):
    import torch.distributed as dist
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = _problem(kind)
    opt = g.SparseOptimizer(device=-1, shard=rank, num_shards=world)
    synth.feed(p, opt)
    opt.setup_cli()
    opt.initialize_optimization()
    opt._ensure_uploaded()
    ctx = opt.context
    assert ctx.build_structure()
    n = g.lib.b200_get_blocks(ctx.handle, 3, None, None, None)
    rows, cols = np.zeros(n, np.int32), np.zeros(n, np.int32)
    g.lib.b200_get_blocks(ctx.handle, 3, rows.ctypes.data, cols.ctypes.data, None)
    d = ctx.dims()
    mine = dict(pattern=(rows.tolist(), cols.tolist()), perm=ctx.block_ordering().tolist(), lnz=ctx.factor_nnz(),
                nl=d["numLandmarks"], ne=d["numEdges"], np=d["numPoses"])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["ba", "ba_expmap"])
def test_landmark_shards_agree_on_the_reduced_system_cpu(kind):
    """every rank must build the identical Hschur pattern / ordering (the all-reduce adds arrays element-wise),
    the shards must partition landmarks and edges, and the pattern must equal the unsharded one"""
    import torch.multiprocessing as mp
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_cpu_worker, args=(r, 2, port, q, kind)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    p = _problem(kind)
    ref = g.SparseOptimizer(device=-1)
    synth.feed(p, ref)
    ref.setup_cli()
    ref.initialize_optimization()
    ref._ensure_uploaded()
    assert ref.context.build_structure()
    d = ref.context.dims()
    assert res[0]["pattern"] == res[1]["pattern"]
    assert res[0]["perm"] == res[1]["perm"] == ref.context.block_ordering().tolist()
    assert res[0]["lnz"] == res[1]["lnz"] == ref.context.factor_nnz()
    assert res[0]["nl"] + res[1]["nl"] == d["numLandmarks"]
    assert res[0]["ne"] + res[1]["ne"] == d["numEdges"]
    assert abs(res[0]["ne"] - res[1]["ne"]) <= 0.1 * d["numEdges"]  # balanced by edge count
    assert res[0]["np"] == res[1]["np"] == d["numPoses"]


def _gpu_worker(rank, world, port, out, native=True):
    import torch
    import torch.distributed as dist
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    from openslam_g2o_b200.distributed import sharded_optimizer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    p = synth.venice_like(40, 6000, seed=12)
    opt = sharded_optimizer(p, rank, world, rank, native=native)
    n = opt.optimize(6)
    chi = [s.chi2 for s in opt.batch_statistics]
    cams = opt.context.estimates(g.VERTEX_CAM, 40)
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(n=n, chi=chi, cams=cams.tolist()))
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("native", [True, False], ids=["nccl_native", "host_callback"])
def test_sharded_bundle_adjustment_matches_single_gpu(native):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_gpu_worker, args=(r, 2, port, q, native)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    p = synth.venice_like(40, 6000, seed=12)
    ref = g.SparseOptimizer(device=0)
    ref.set_algorithm("lm_fix6_3")
    synth.feed(p, ref)
    ref.setup_cli()
    ref.initialize_optimization()
    n = ref.optimize(6)
    chi = np.array([s.chi2 for s in ref.batch_statistics])
    cams = ref.context.estimates(g.VERTEX_CAM, 40)
    for r in res:
        assert r["n"] == n
        assert np.abs(np.array(r["chi"]) - chi).max() <= 1e-9 * chi.max()
        assert np.abs(np.array(r["cams"]) - cams).max() <= 1e-8 * np.abs(cams).max()
    assert res[0]["chi"] == res[1]["chi"]  # ranks stay in lock-step bit for bit


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_download_scatters_only_the_owned_landmarks(world):
    """b200_graph_download on a landmark-sharded upload: the context holds only this shard's landmarks in local row
    order - every owned landmark must land in ITS vertex, the others keep their host estimates (host-only context:
    the estimates come back as ingested, so any mis-scatter shows up as a changed vertex)"""
    import openslam_g2o_b200 as g
    from openslam_g2o_b200 import synth
    p = synth.venice_like(20, 700, seed=4)
    owned_total = 0
    for rank in range(world):
        opt = g.SparseOptimizer(device=-1, shard=rank, num_shards=world)
        synth.feed(p, opt)
        opt.setup_cli()
        opt.initialize_optimization()
        opt._ensure_uploaded()
        assert opt.context.build_structure()
        owned_total += opt.context.dims()["numLandmarks"]
        before = {int(i): opt.vertex_estimate(int(i)) for i in p["point_ids"]}
        opt.sync_estimates()
        for i in p["point_ids"]:
            assert np.array_equal(opt.vertex_estimate(int(i)), before[int(i)]), (rank, int(i))
        for i in p["cam_ids"][:5]:
            assert np.allclose(opt.vertex_estimate(int(i))[:7], p["cam_payload"][int(i)][:7])
    assert owned_total == len(p["point_ids"])
