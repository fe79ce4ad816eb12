"""Communicator bootstrap for landmark-sharded bundle adjustment.

One process per GPU.  Landmarks (and their observations) are split over the ranks, cameras are replicated.  The data
path is native: the C++ host issues ncclAllReduce itself on the solver's stream (csrc/nccl_comm.cpp, captured in the
trial CUDA graph) - two per LM trial: [Hschur | bschur | chi2] before the factorisation, 2 doubles after the update.
Python only carries the NCCL unique id from rank 0 to the other ranks (`init_native_comm`, through the already
initialised torch.distributed process group - any out-of-band channel would do).

`make_allreduce` is the older host-callback route (b200_set_allreduce): the same protocol, every collective a Python
round trip into torch.distributed and no CUDA graphs - kept for hosts that own their communicator.
"""
import sys


def init_native_comm(ctx, rank, world_size):
    """collective: rank 0 creates the NCCL id, everybody joins (ncclCommInitRank on ctx's device)"""
    import torch.distributed as dist
    from .optimizer import comm_unique_id
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world_size)


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


def sharded_optimizer(problem, rank, world_size, local_rank, algorithm="lm_fix6_3", native=True):
    """SparseOptimizer over this rank's landmark shard, wired to the process group (must be initialised)"""
    from . import SparseOptimizer, synth
    opt = SparseOptimizer(device=local_rank, shard=rank, num_shards=world_size)
    opt.set_algorithm(algorithm)
    synth.feed(problem, opt)
    opt.setup_cli()
    opt.initialize_optimization()
    opt._ensure_uploaded()
    if world_size > 1:
        if native:
            init_native_comm(opt.context, rank, world_size)
        else:
            opt.context.set_allreduce(make_allreduce(opt.context, local_rank), rank, world_size)
    return opt
