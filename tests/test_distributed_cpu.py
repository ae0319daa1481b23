"""world_size-2 gloo tests of the multi-GPU host logic: shard the batch, sum the
calibration histograms over ranks, gather the logits rank-major.  Runs on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from term_quantization_b200 import inference, tr_layer
        torch.manual_seed(0)
        batch = torch.randn(10, 4)                       # the whole job's batch, same on all ranks
        lo, hi = inference.shard_bounds(10, world, rank)
        local = batch[lo:hi] * 2.0                        # stands in for model(shard)
        full = inference.gather_logits(local)
        ok_gather = torch.equal(full, batch * 2.0)

        q = torch.nn.Sequential(tr_layer.LinearQuantize(8, 3), tr_layer.LinearQuantize(8, 3))
        for i, m in enumerate(q):
            m.hist_bins[rank + i] = float(rank + 1)      # rank-specific calibration counts
        inference.allreduce_histograms(q)
        ok_hist = all(float(m.hist_bins.sum()) == 3.0 and float(m.hist_bins[i]) == 1.0 and
                      float(m.hist_bins[i + 1]) == 2.0 for i, m in enumerate(q))
        # one gather for a whole run (gather="end"): same rank-major order per step as per-step gathers
        steps = [batch[lo:hi] * float(k + 1) for k in range(3)]
        all_steps = inference.gather_steps(steps)
        ok_steps = all_steps.shape == (3, 10, 4) and all(torch.equal(all_steps[k], batch * float(k + 1)) for k in range(3))
        ret[rank] = (ok_gather, ok_hist and ok_steps)
    finally:
        dist.destroy_process_group()


def test_world2_gather_and_histogram_allreduce():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: (True, True), 1: (True, True)}
