"""N-rank NCCL check of the batch-sharded path (run under torchrun on an N-GPU box):
every rank computes the Chamfer-L1 loss of ITS shard through parallel.sharded_chamfer (local
kernels + one 16-byte all-reduce) and the result must equal the single-GPU loss of the whole batch
(rank 0 computes it with the unsharded module) within fp32 summation-order tolerance; gradients of
the local clouds must equal the corresponding rows of the unsharded gradients to 1e-5
(no collective in backward).  Group / FPS need no exchange: checked to be shard-invariant.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/check_multigpu.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "iccv2025-upp_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import upp_b200
    from upp_b200 import parallel

    g = torch.Generator().manual_seed(7)
    B = 4 * world + 1  # uneven on purpose: remainders go to the low ranks
    a = torch.rand(B, 700, 3, generator=g)
    b = torch.rand(B, 900, 3, generator=g)
    lo, hi = parallel.shard_bounds(B, rank, world)
    for kind, mod in (("l1", upp_b200.ChamferDistanceL1()), ("l2", upp_b200.ChamferDistanceL2())):
        al = a[lo:hi].to(dev).requires_grad_(True)
        bl = b[lo:hi].to(dev).requires_grad_(True)
        loss = parallel.sharded_chamfer(al, bl, kind, n_global_clouds=B)
        loss.backward()
        af = a.to(dev).requires_grad_(True)
        bf = b.to(dev).requires_grad_(True)
        full = mod(af, bf)
        full.backward()
        rel = abs(loss.item() - full.item()) / abs(full.item())
        assert rel <= 1e-5, (kind, loss.item(), full.item())
        # same kernels, but the scalar chain differs (1/(4 n sqrt d) here vs autograd's mean -> sqrt): 1e-5 rel
        for got, want in ((al.grad, af.grad[lo:hi]), (bl.grad, bf.grad[lo:hi])):
            torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5 * float(want.abs().max()))
    # fused peer-memory all-reduce (upp_chamfer_fwd_sharded_f32) against the NCCL path: same loss to fp32 summation
    # order, the SAME BITS on every rank, stable over many calls (slot parity reuse) and under CUDA-graph replay
    try:
        peers = parallel.PeerExchange()
    except RuntimeError as ex:
        peers = None
        if rank == 0:
            print("peer exchange unavailable, NCCL path only:", str(ex)[:300])
    if peers is not None:
        al, bl = a[lo:hi].to(dev), b[lo:hi].to(dev)
        for it in range(5):
            for kind in ("l1", "l2"):
                x = (al + 0.01 * it).requires_grad_(True)
                y = bl.clone().requires_grad_(True)
                fused = parallel.sharded_chamfer(x, y, kind, n_global_clouds=B, peers=peers)
                fused.backward()
                gx = x.grad.clone()
                x2 = (al + 0.01 * it).requires_grad_(True)
                nccl = parallel.sharded_chamfer(x2, bl.clone().requires_grad_(True), kind, n_global_clouds=B)
                nccl.backward()
                assert torch.isfinite(fused), "peer wait timed out"
                assert abs(fused.item() - nccl.item()) <= 1e-6 * abs(nccl.item()), (kind, it, fused.item(), nccl.item())
                assert torch.equal(gx, x2.grad)  # same local kernels, deterministic backward: bit-equal
                bits = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(world)]
                dist.all_gather(bits, fused.detach().view(torch.int32).reshape(1))
                assert all(torch.equal(bits[0], v) for v in bits), "fused all-reduce must be bit-identical across ranks"
        # CUDA graph: the exchange (device-side sequence counter) replays
        sx, sy = al.clone(), bl.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                upp_b200.ops.chamfer_forward_sharded(sx, sy, peers)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = upp_b200.ops.chamfer_forward_sharded(sx, sy, peers)
        want = None
        for it in range(6):
            sx.copy_(al + 0.02 * it)
            graph.replay()
            torch.cuda.synchronize()
            ref = upp_b200.ops.chamfer_forward(sx, sy, want_sums=True)[4]
            dist.all_reduce(ref)
            torch.testing.assert_close(out[4], ref, rtol=1e-6, atol=0)
        # deferred exchange: send inside the Chamfer kernels, wait + sum in a later launch
        for it in range(4):
            sx.copy_(al + 0.03 * it)
            loc = upp_b200.ops.chamfer_forward_sharded(sx, sy, peers, defer=True)[4].clone()
            glob = upp_b200.ops.peer_allreduce_finish(peers, dev)
            ref = loc.clone()
            dist.all_reduce(ref)
            torch.testing.assert_close(glob, ref, rtol=1e-6, atol=0)
        if rank == 0:
            print(f"fused peer all-reduce ok via {peers.how}: == NCCL (1e-6), bit-identical across ranks, graph replay x6, deferred finish x4")
        peers.check()  # no exchange timed out
    # gradient statistics (sum ||grad||^2 over the GLOBAL batch, exchanged inside the backward kernel) against
    # torch.nn.utils.clip_grad_norm_'s total norm of the unsharded coordinate gradients (tools/runner_module.py:204);
    # DDP-compatible scaling: grad_scale="ddp" makes the rank-AVERAGED gradient equal the unsharded one;
    # an EMPTY shard (global batch smaller than the world) takes part in the exchange with zeros
    for use_peers in ((peers, None) if peers is not None else (None,)):
        for kind, mod in (("l2", upp_b200.ChamferDistanceL2()), ("l1", upp_b200.ChamferDistanceL1())):
            af = (a.to(dev) + 0.003).requires_grad_(True)   # no exact-zero distances: L1 gradients finite
            bf = b.to(dev).requires_grad_(True)
            mod(af, bf).backward()
            want_norm = torch.nn.utils.clip_grad_norm_([af, bf], max_norm=1e30)
            al = (a[lo:hi].to(dev) + 0.003).requires_grad_(True)
            bl = b[lo:hi].to(dev).requires_grad_(True)
            st = parallel.GradStats()
            loss = parallel.sharded_chamfer(al, bl, kind, n_global_clouds=B, peers=use_peers, stats=st)
            loss.backward()
            got_norm = st.total_norm()
            assert abs(got_norm.item() - want_norm.item()) <= 1e-4 * want_norm.item(), (kind, got_norm.item(), want_norm.item())
            bits = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in range(world)]
            dist.all_gather(bits, st.sq_norm.view(torch.int32).clone())
            assert all(torch.equal(bits[0], v) for v in bits) or use_peers is None, "fused statistics must be bit-identical across ranks"
            again = parallel.GradStats()
            al2 = (a[lo:hi].to(dev) + 0.003).requires_grad_(True)
            parallel.sharded_chamfer(al2, b[lo:hi].to(dev), kind, n_global_clouds=B, peers=use_peers, stats=again).backward()
            assert torch.equal(again.sq_norm, st.sq_norm) and torch.equal(al2.grad, al.grad), "deterministic backward + statistics"
            # DDP semantics: average over ranks of the 'ddp'-scaled local gradients == unsharded gradient rows
            al3 = (a[lo:hi].to(dev) + 0.003).requires_grad_(True)
            parallel.sharded_chamfer(al3, b[lo:hi].to(dev), kind, n_global_clouds=B, peers=use_peers, grad_scale="ddp").backward()
            torch.testing.assert_close(al3.grad / world, al.grad, rtol=1e-6, atol=0)
            torch.testing.assert_close(al.grad, af.grad[lo:hi], rtol=1e-4, atol=1e-6 * float(af.grad.abs().max()))
        # empty shard: 1 cloud in total, world ranks
        lo1, hi1 = parallel.shard_bounds(1, rank, world)
        e1 = a[:1][lo1:hi1].to(dev).requires_grad_(True)
        e2 = b[:1][lo1:hi1].to(dev)
        st = parallel.GradStats()
        loss = parallel.sharded_chamfer(e1, e2, "l2", n_global_clouds=1, peers=use_peers, stats=st)
        loss.backward()
        want = upp_b200.ChamferDistanceL2()(a[:1].to(dev), b[:1].to(dev))
        assert abs(loss.item() - want.item()) <= 1e-5 * want.item(), ("empty shard", rank, loss.item(), want.item())
        assert e1.grad.shape[0] == hi1 - lo1 and torch.isfinite(st.sq_norm).all()
    if peers is not None:
        peers.check()
    # Group is per cloud: a shard's result equals the same rows of the full batch
    x = (torch.rand(B, 1024, 3, generator=g) * 2 - 1).to(dev)
    nb_f, ce_f = upp_b200.Group(64, 32)(x)
    nb_l, ce_l = upp_b200.Group(64, 32)(x[lo:hi].contiguous())
    assert torch.equal(nb_f[lo:hi], nb_l) and torch.equal(ce_f[lo:hi], ce_l)
    ok = torch.ones(1, device=dev)
    dist.all_reduce(ok)
    if rank == 0:
        print(f"multigpu ok: {int(ok.item())}/{world} ranks, sharded Chamfer L1/L2 loss and grads == unsharded (1e-5), gradient statistics == "
              f"clip_grad_norm_ total norm, ddp scaling, empty shard, Group shard-invariant (bit-equal)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
