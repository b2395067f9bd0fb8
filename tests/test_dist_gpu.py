"""Two-GPU parity of the row-sharded Cox forward/backward: both exchanges -- the one fused into the
forward kernel over NVLink peer memory and the NCCL all-reduce -- must reproduce the single-GPU loss BIT FOR BIT
(per-bin sums are integers) and the single-GPU gradient on every rank's rows.  Skipped on a one-GPU box."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_q):
    import ctypes

    import torch.distributed as dist
    from multimodal_survival_prediction_b200 import _lib as L
    from multimodal_survival_prediction_b200 import dist as bd
    from multimodal_survival_prediction_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {}
    try:
        a, b = bd.shard_bounds(n, rank, world)
        for exchange in ("peer", "nccl"):
            op = bd.ShardedCoxBinned(b - a, dev, nbins=4096, ties="efron", exchange=exchange)
            assert op.exchange == exchange
            for it in range(4):                    # several epochs: exercises the double-buffered slots
                lh, ev, t = synth.cohort(n, 100 + it)
                x, e, tt = lh[a:b].to(dev), ev[a:b].to(dev), t[a:b].to(dev)
                grad = torch.empty(b - a, dtype=torch.float32, device=dev)
                loss = op.forward(x, tt, e)
                op.backward(x, tt, e, grad)
                # single-GPU result for the whole cohort on this rank's device, same kernels
                xf, ef, tf = lh.to(dev), ev.to(dev), t.to(dev)
                lib = L.load()
                sb = lib.b200surv_cox_state_bytes(n, 1, L.COX_BINNED, 4096)
                wb = lib.b200surv_cox_workspace_bytes(n, 1, L.COX_BINNED, 4096)
                state = torch.empty(sb, dtype=torch.uint8, device=dev)
                ws = torch.empty(wb, dtype=torch.uint8, device=dev)
                l1 = torch.empty(1, dtype=torch.float32, device=dev)
                L.check(lib.b200surv_cox_fwd(L.ptr(xf), L.ptr(tf), L.ptr(ef), None, n, 1, 2, 0, L.COX_BINNED, 4096,
                                             ctypes.c_float(0.0), L.ptr(l1), L.ptr(state), sb, L.ptr(ws), wb,
                                             L.stream_ptr(dev)), "fwd")
                gf = torch.empty(n, dtype=torch.float32, device=dev)
                one = torch.ones(1, dtype=torch.float32, device=dev)
                L.check(lib.b200surv_cox_bwd(L.ptr(one), L.ptr(state), sb, L.ptr(xf), L.ptr(tf), L.ptr(ef), None, n, 1,
                                             L.COX_BINNED, 4096, L.ptr(gf), L.stream_ptr(dev)), "bwd")
                torch.cuda.synchronize()
                res[(exchange, it)] = (loss.item(), l1.item(), bool(torch.equal(grad, gf[a:b])))
            if op.peers is not None:
                dist.barrier()
                op.peers.close()
        out_q.put((rank, res, None))
    except Exception as ex:  # noqa: BLE001 -- reported to the parent
        out_q.put((rank, res, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_sharded_cox_two_gpus_bit_identical():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    n, world = (1 << 20) + 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, res, err in got:
        assert err is None, f"rank {rank}: {err}"
        assert len(res) == 8
        for key, (loss, loss_single, grad_equal) in res.items():
            assert loss == loss_single, (rank, key, loss, loss_single)   # bit-identical fp32 loss
            assert grad_equal, (rank, key)


def _worker_sorted(rank, world, port, n, out_q):
    import numpy as np
    import torch.distributed as dist
    import multimodal_survival_prediction_b200 as pkg
    from multimodal_survival_prediction_b200 import dist as bd
    from multimodal_survival_prediction_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {}
    try:
        for it, few in enumerate((True, False)):     # continuous times; integer days (tie groups across the shard edge)
            lh, ev, t = synth.cohort(n, 200 + it, few_ties=few)
            order = torch.argsort(t, stable=True)
            a, b = bd.shard_bounds(n, rank, world)
            idx = order[a:b][torch.randperm(b - a, generator=torch.Generator().manual_seed(rank))]
            op = bd.ShardedCoxSorted(b - a, dev, ties="efron")
            x, tt, e = lh[idx].to(dev), t[idx].to(dev), ev[idx].to(dev)
            grad = torch.empty(b - a, dtype=torch.float32, device=dev)
            for _ in range(2):                       # twice: the workspace and the records are reused
                loss = op.forward(x, tt, e)
                op.backward(grad)
            assert op.check() == 0
            xf = lh.to(dev).requires_grad_(True)
            l1 = pkg.neg_partial_log_likelihood(xf, ev.to(dev), t.to(dev), "efron", mode="sorted")
            l1.backward()
            torch.cuda.synchronize()
            gref = xf.grad[idx.to(dev)]
            res[it] = (float(loss), float(l1), float((grad - gref).abs().max()), float(xf.grad.abs().max()))
        out_q.put((rank, res, None))
    except Exception as ex:  # noqa: BLE001 -- reported to the parent
        out_q.put((rank, res, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_sharded_cox_sorted_two_gpus_time_range_shards():
    """north_star's multi-GPU Cox: time-range shards, NCCL all-gather of the per-shard boundary records, exact carry-in.
    The loss (same on both ranks) and every rank's gradient rows equal the single-GPU SORTED result to 2e-6 (fp64 sums
    in another order)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    n, world = (1 << 19) + 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sorted, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, res, err in got:
        assert err is None, f"rank {rank}: {err}"
        assert len(res) == 2
        for key, (loss, loss_single, gerr, gmax) in res.items():
            assert abs(loss - loss_single) <= 2e-6 * abs(loss_single), (rank, key, loss, loss_single)
            assert gerr <= 2e-6 * gmax, (rank, key, gerr, gmax)
    assert got[0][1][0][0] == got[1][1][0][0]        # both ranks hold the same loss bits


def _worker_rowblock(rank, world, port, n, out_q):
    import torch.distributed as dist
    import multimodal_survival_prediction_b200 as pkg
    from multimodal_survival_prediction_b200 import dist as bd
    from multimodal_survival_prediction_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {}
    try:
        for it, few in enumerate((True, False)):
            lh, ev, t = synth.cohort(n, 300 + it, few_ties=few)
            a, b = bd.shard_bounds(n, rank, world)                  # ROW blocks: every rank holds times of the whole range
            op = bd.RowBlockCoxSorted(b - a, dev).plan(t[a:b].to(dev), ev[a:b].to(dev))
            grad = torch.empty(b - a, dtype=torch.float32, device=dev)
            for step in range(2):                                   # the plan is reused: only log_hz travels per step
                x = (lh[a:b] + 0.25 * step).to(dev)
                loss = op.forward(x)
                op.backward(grad)
            assert op.check() == 0
            xf = (lh + 0.25).to(dev).requires_grad_(True)
            l1 = pkg.neg_partial_log_likelihood(xf, ev.to(dev), t.to(dev), "efron", mode="sorted")
            l1.backward()
            torch.cuda.synchronize()
            res[it] = (float(loss), float(l1.detach()), float((grad - xf.grad[a:b]).abs().max()), float(xf.grad.abs().max()),
                       op.n_recv)
        out_q.put((rank, res, None))
    except Exception as ex:  # noqa: BLE001 -- reported to the parent
        out_q.put((rank, res, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_row_block_cox_sorted_two_gpus():
    """Patients sharded by ROW BLOCK (north_star): sample sort on time across the ranks (b200surv_route_rows + all-to-all),
    then the time-range shard phases.  Loss and each rank's gradient rows equal the single-GPU SORTED result (2e-6)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    n, world = (1 << 19) + 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_rowblock, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    total = {0: 0, 1: 0}
    for rank, res, err in got:
        assert err is None, f"rank {rank}: {err}"
        for key, (loss, loss_single, gerr, gmax, n_recv) in res.items():
            assert abs(loss - loss_single) <= 2e-6 * abs(loss_single), (rank, key, loss, loss_single)
            assert gerr <= 2e-6 * gmax, (rank, key, gerr, gmax)
            assert 0.4 * n <= n_recv <= 0.6 * n, (rank, key, n_recv)        # the sample sort balances the ranks
            total[key] += n_recv
    assert total[0] == n and total[1] == n
