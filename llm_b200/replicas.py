"""Multi-GPU host plumbing for the replica layout: one process per GPU, each with its own model and KV cache, no
data-path collective.  torch.distributed is used for exactly two things: the barrier that brackets the timed region and the
MAX-over-ranks of the device-timed milliseconds.  Backend "nccl" on the GPU box, "gloo" in the CPU tests."""
import os


class Replicas:
    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        self.backend = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self.backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            if self.backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(self.backend)
            self.dist = dist

    def barrier(self):
        if self.dist is None:
            return
        if self.backend == "nccl":
            import torch
            self.dist.barrier(device_ids=[self.local_rank])
            torch.cuda.synchronize()
        else:
            self.dist.barrier()

    def max_over_ranks(self, values):
        """element-wise MAX of a list of floats over all ranks (every rank gets the result)"""
        if self.dist is None:
            return list(values)
        import torch
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = torch.tensor(list(values), dtype=torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def whole_job_rate(self, units_per_rank, ms_max):
        """weak scaling: every rank processed `units_per_rank` units; the job took the slowest rank's time"""
        return self.world * units_per_rank / (ms_max * 1e-3)

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None
