"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a replicated batch, each runs its slice, and the
summed loss is all-reduced.  The per-slice computation is stood in for by the oracle (tests may use it); what is under
test is tf_seq2seq_losses_b200.sharding, which the N-GPU benchmark uses unchanged with the CUDA entry point."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tf_seq2seq_losses_b200.sharding import shard_bounds


def test_shard_bounds_cover_the_batch():
    for batch in (0, 1, 7, 256, 2048):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ctc_oracle as orc
    from tests.ref_cases import random_inputs
    from tf_seq2seq_losses_b200.sharding import sharded_loss_and_grad

    logits, labels, ll, tl = random_inputs(7, 12, 6, 4, seed=3)
    ll[2], tl[2] = 4, 2        # one infeasible sample: excluded from the reduced sum

    def loss_grad_fn(lab, x, l1, l2, blank):
        loss, grad, _ = orc.loss_and_grad_logits(lab.numpy(), x.numpy(), l1.numpy(), l2.numpy(), blank, orc.CLASSIC)
        return torch.tensor(loss, dtype=torch.float32), torch.tensor(grad, dtype=torch.float32)

    loss, grad, (b0, b1), total = sharded_loss_and_grad(loss_grad_fn, torch.tensor(labels), torch.tensor(logits),
                                                         torch.tensor(ll), torch.tensor(tl), 0)
    torch.save({"loss": loss, "span": (b0, b1), "total": total}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import ctc_oracle as orc
    from tests.ref_cases import random_inputs
    logits, labels, ll, tl = random_inputs(7, 12, 6, 4, seed=3)
    ll[2], tl[2] = 4, 2
    want, _, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, orc.CLASSIC)
    parts = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    got = np.concatenate([p["loss"].numpy() for p in parts])
    assert parts[0]["span"] == (0, 4) and parts[1]["span"] == (4, 7)
    assert np.isinf(got[2]) and np.isinf(want[2])
    fin = np.isfinite(want)
    assert np.allclose(got[fin], want[fin], rtol=1e-6)
    for p in parts:     # every rank holds the same all-reduced total over the feasible samples
        assert abs(float(p["total"]) - float(want[fin].sum())) < 1e-3
