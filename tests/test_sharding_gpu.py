"""Multi-GPU path on real devices: two NCCL ranks shard a replicated batch through tf_seq2seq_losses_b200.sharding, each
runs its slice through the CUDA library, and the summed loss is all-reduced over NVLink.  Needs two GPUs (`gpurun --gpus 2`);
on a one-GPU box the test is skipped (the same host logic runs on CPU under gloo in tests/test_sharding_cpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs():
    from tests.ref_cases import random_inputs
    logits, labels, ll, tl = random_inputs(9, 40, 64, 8, seed=13)
    ll[2], tl[2] = 8, 3        # one infeasible sample: excluded from the reduced sum
    return logits, labels, ll, tl


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from tf_seq2seq_losses_b200 import _lib
    from tf_seq2seq_losses_b200.sharding import sharded_loss_and_grad

    logits, labels, ll, tl = (torch.as_tensor(a).to(dev) for a in _inputs())

    def loss_grad_fn(lab, x, l1, l2, blank):
        x = x.contiguous()
        desc = _lib.make_desc(x, lab, blank, _lib.CLASSIC, 9)
        loss, grad, _ = _lib.loss_grad(desc, x, lab.contiguous(), l1.contiguous(), l2.contiguous())
        return loss, grad

    loss, grad, (b0, b1), total = sharded_loss_and_grad(loss_grad_fn, labels, logits, ll, tl, 0)
    torch.save({"loss": loss.cpu(), "grad": grad.cpu(), "span": (b0, b1), "total": total.cpu()},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_nccl_sharding(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import ctc_oracle as orc
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    logits, labels, ll, tl = _inputs()
    want, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, orc.CLASSIC)
    parts = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    assert parts[0]["span"] == (0, 5) and parts[1]["span"] == (5, 9)
    got = np.concatenate([p["loss"].numpy() for p in parts])
    got_grad = np.concatenate([p["grad"].numpy() for p in parts])
    assert np.isinf(got[2]) and np.isinf(want[2])
    fin = np.isfinite(want)
    assert np.allclose(got[fin], want[fin], rtol=1e-5)
    want_grad[~fin] = 0.0
    assert np.max(np.abs(got_grad - want_grad)) <= 5e-5
    for p in parts:     # every rank holds the same all-reduced total over the feasible samples
        assert abs(float(p["total"]) - float(want[fin].sum())) < 1e-2
