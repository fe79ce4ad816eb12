"""torch.distributed plumbing for landmark-sharded bundle adjustment.

One process per GPU.  Landmarks (and their observations) are split over the ranks, cameras are replicated; the only
data-path exchange is the all-reduce of [Hschur | bschur] per LM trial (+ two tiny ones: partial Hpp/b_p after
linearisation, chi2/scale scalars).  The C library calls back into `make_allreduce()` with a raw device pointer; the
callback wraps it as a tensor (no copy) and issues the NCCL collective ordered on the solver's CUDA stream.
"""
import sys


class _DevicePointer:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def make_allreduce(ctx, local_rank):
    """returns a python callable with the b200_allreduce_fn signature, bound to ctx's stream"""
    import torch
    import torch.distributed as dist
    device = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)

    views = {}  # (ptr, count) -> tensor view: the library reduces the same few device buffers every iteration

    def allreduce(ptr, count, op, _stream, _user):
        try:
            t = views.get((ptr, count))
            if t is None:
                t = views[(ptr, count)] = torch.as_tensor(_DevicePointer(ptr, count), device=device)
            with torch.cuda.stream(stream):
                dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM)
            return 0
        except Exception as e:  # noqa: BLE001 - must not propagate through the C frame
            print("g2o_b200 all-reduce callback failed: %r" % (e,), file=sys.stderr)
            return 1
    return allreduce


def sharded_optimizer(problem, rank, world_size, local_rank, algorithm="lm_fix6_3"):
    """SparseOptimizer over this rank's landmark shard, wired to the process group (must be initialised)"""
    from . import SparseOptimizer, synth
    opt = SparseOptimizer(device=local_rank, shard=rank, num_shards=world_size)
    opt.set_algorithm(algorithm)
    synth.feed(problem, opt)
    opt.setup_cli()
    opt.initialize_optimization()
    opt._ensure_uploaded()
    if world_size > 1:
        opt.context.set_allreduce(make_allreduce(opt.context, local_rank), rank, world_size)
    return opt
