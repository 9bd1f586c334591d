"""CPU, world_size 2, gloo: the host-side logic of the N > 1 path (mpvae_b200.train / .infer).

The CUDA loss cannot run here, so the oracle's `compute_loss` is injected as `loss_fn` (tests may use the
oracle); what is under test is the data-parallel machinery around it: row sharding, the flat gradient bucket,
the early g_R all-reduce hook, weighting of unequal shards, identical clip + Adam on every rank, the float64
r_sqrt_sigma round trip, and the inference gather.  Criterion: two ranks on half batches == one process on the
full batch (same noise), parameter for parameter."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_args():
    return SimpleNamespace(feature_dim=12, label_dim=9, latent_dim=6, z_dim=5, keep_prob=0.0, scale_coeff=1.0,
                           residue_sigma="", n_train_sample=4, n_test_sample=6, mode="train", nll_coeff=0.5,
                           c_coeff=10.0, batch_size=8)


def make_data(n=23, seed=0):
    rng = np.random.RandomState(seed)
    a = make_args()
    x = torch.from_numpy(rng.standard_normal((n, a.feature_dim)).astype(np.float32))
    y = (rng.uniform(size=(n, a.label_dim)) < 0.3).astype(np.float32)
    y[:, 0], y[:, 1] = 1.0, 0.0
    return x, torch.from_numpy(y)


def build(args, sgd=False):
    sys.path.insert(0, ROOT)
    from mpvae_b200.mpvae import VAE
    from mpvae_b200.train import DataParallelStep
    from oracle import probit_elbo_oracle as orc
    np.random.seed(4)
    torch.manual_seed(0)
    model = VAE(args)
    # (SGD for the 2-rank comparison: Adam turns 1e-9 gradient noise on dead units into +-lr parameter moves)
    opt = torch.optim.SGD(model.parameters(), lr=0.05, weight_decay=1e-5) if sgd else \
        torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.StepLR(opt, 3, 0.5)

    def loss_fn(y, fe_out, fe_mu, fe_lv, fx_out, fx_mu, fx_lv, r, a, noise=None):
        return orc.compute_loss(y, fe_out, fe_mu, fe_lv, fx_out, fx_mu, fx_lv, r, a, noise=noise, ranking="factorised")

    return model, DataParallelStep(model, opt, sched, args, clip_norm=100.0, loss_fn=loss_fn), loss_fn


def noise_for(step_i, S, n, Z):
    g = torch.Generator().manual_seed(1000 + step_i)
    return torch.randn(S, n, Z, generator=g)


def run_steps(n_steps=3, batch=11):
    args = make_args()
    model, stepper, loss_fn = build(args, sgd=True)
    x, y = make_data()
    # keep_prob=0 (dropout off) but the reparameterisation still draws randn_like: make it rank-independent by
    # fixing the generator per step (both ranks draw the SAME eps for row i only if they see the same rows), so
    # instead remove the stochastic part: eval-mode dropout + zero logvar heads
    with torch.no_grad():
        for head in (model.fe_logvar, model.fx_logvar):
            head.weight.zero_(); head.bias.fill_(-30.0)       # std = exp(-15): eps contributes < 1e-6
    outs = []
    for i in range(n_steps):
        idx = torch.arange(i * 6, i * 6 + batch) % x.shape[0]
        noise = noise_for(i, args.n_train_sample, batch, args.z_dim)
        outs.append(stepper.step(y[idx], x[idx], noise=noise))
    return model, outs


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        if ret.get("mode") == "one_row":
            # a ragged last batch with fewer rows than ranks: rank 1 owns no row and must still join every collective
            model, outs = run_steps(n_steps=2, batch=1)
            ret[rank] = ({k: v.clone() for k, v in model.state_dict().items()},
                         [tuple(float(t) for t in o[:6]) for o in outs], [float(o.grad_norm) for o in outs], None)
            return
        model, outs = run_steps()
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        # inference: sharded scoring + gather
        from mpvae_b200.infer import predict_proba
        from oracle import probit_elbo_oracle as orc
        x, y = make_data()
        args = make_args()

        def infer_loss(yy, *rest, **kw):
            a = rest[-1]
            g = torch.Generator().manual_seed(77)
            full = torch.randn(a.n_test_sample, x.shape[0], a.z_dim, generator=g)
            return orc.compute_loss(yy, *rest, noise=full[:, a.dp_row0:a.dp_row0 + yy.shape[0]], ranking="factorised")

        probs, _ = predict_proba(model, x, y, args, batch_size=5, loss_fn=infer_loss)
        ret[rank] = (sd, [tuple(float(t) for t in o[:6]) for o in outs], [float(o.grad_norm) for o in outs], probs)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_equal_one_process():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    # single process, full batch
    sys.path.insert(0, ROOT)
    model, outs = run_steps()
    ref_sd = model.state_dict()
    for rank in range(2):
        sd, scal, norms, probs = ret[rank]
        assert sd["r_sqrt_sigma"].dtype == torch.float64
        for k in ref_sd:
            assert torch.allclose(sd[k].double(), ref_sd[k].double(), rtol=2e-4, atol=2e-6), (rank, k)
        for got, want in zip(scal, outs):
            for a, b in zip(got, want[:6]):
                assert abs(a - float(b)) <= 2e-5 * max(1.0, abs(float(b))), (rank, a, float(b))
        for a, b in zip(norms, outs):
            assert abs(a - float(b.grad_norm)) <= 1e-4 * max(1.0, float(b.grad_norm))
    # both ranks hold identical replicas and identical gathered predictions
    for k in ref_sd:
        assert torch.equal(ret[0][0][k], ret[1][0][k]), k
    assert torch.equal(ret[0][3], ret[1][3])
    assert ret[0][3].shape == (23, 9)


@pytest.mark.timeout(300)
def test_fewer_rows_than_ranks():
    """N % batch in [1, world) leaves a rank with an empty shard.  The collective sequence must not depend
    on local data (no hang), the empty rank contributes zero gradients and zero-weighted scalars (no NaN), and the result
    equals the single-process step on that one row."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    ret["mode"] = "one_row"
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    model, outs = run_steps(n_steps=2, batch=1)
    ref_sd = model.state_dict()
    for rank in range(2):
        sd, scal, norms, _ = ret[rank]
        for k in ref_sd:
            assert torch.allclose(sd[k].double(), ref_sd[k].double(), rtol=2e-4, atol=2e-6), (rank, k)
        for got, want in zip(scal, outs):
            for a, b in zip(got, want[:6]):
                assert np.isfinite(a) and abs(a - float(b)) <= 2e-5 * max(1.0, abs(float(b))), (rank, a, float(b))
        for a, b in zip(norms, outs):
            assert abs(a - float(b.grad_norm)) <= 1e-4 * max(1.0, float(b.grad_norm))


def test_shard_rows_cover_everything():
    from mpvae_b200.train import shard_rows
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_single_process_step_matches_plain_autograd():
    """world_size 1: DataParallelStep == the literal loop body of train.py:103-129."""
    from oracle import probit_elbo_oracle as orc
    args = make_args()
    model, stepper, loss_fn = build(args)
    x, y = make_data()
    np.random.seed(4); torch.manual_seed(0)
    from mpvae_b200.mpvae import VAE
    twin = VAE(args)
    twin.load_state_dict(model.state_dict())
    opt = torch.optim.Adam(twin.parameters(), lr=1e-2, weight_decay=1e-5)
    noise = noise_for(0, args.n_train_sample, 11, args.z_dim)
    torch.manual_seed(5)
    out = stepper.step(y[:11], x[:11], noise=noise)
    torch.manual_seed(5)
    opt.zero_grad()
    o = twin(y[:11], x[:11])
    terms = orc.compute_loss(y[:11], *o[:3], *o[3:], twin.r_sqrt_sigma, args, noise=noise, ranking="factorised")
    terms[0].backward()
    torch.nn.utils.clip_grad_norm_(twin.parameters(), 100.0)
    opt.step()
    assert abs(float(out.total_loss) - float(terms[0])) < 1e-5
    for (k, a), (_, b) in zip(model.state_dict().items(), twin.state_dict().items()):
        assert torch.allclose(a.double(), b.double(), rtol=1e-5, atol=1e-7), k


def test_label_bits_round_trip():
    """pack_labels / unpack_labels (the host format of the label matrix in bench.py's e2e leg): lossless for any width."""
    import numpy as np
    import pytest
    import torch
    from mpvae_b200.train import pack_labels, unpack_labels
    rng = np.random.RandomState(3)
    for B, L in ((1, 1), (5, 8), (7, 14), (16, 3993), (3, 81)):
        y = (rng.uniform(size=(B, L)) < 0.3).astype(np.float32)
        bits = pack_labels(y)
        assert bits.dtype == torch.uint8 and tuple(bits.shape) == (B, (L + 7) // 8)
        back = unpack_labels(bits, L)
        assert back.dtype == torch.float32 and torch.equal(back, torch.from_numpy(y))
    with pytest.raises(ValueError):
        pack_labels(np.array([[0.0, 0.5]]))


def test_grad_bucket_layout():
    """[g_R | pad | MLP gradients | pad | 8 scalar slots]: segments start at multiples of four floats (16-byte accesses of
    the peer-memory exchange), `grads` excludes the scalar slots, padding stays zero, external storage is accepted."""
    import torch
    from mpvae_b200.train import GradBucket
    r = torch.zeros(7, 9, requires_grad=True)                     # 63 floats: an odd segment like eurlex's 3993 x 3993
    ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(2)), torch.nn.Parameter(torch.zeros(1), requires_grad=False)]
    for alloc in (None, lambda n: torch.full((n,), 7.0)):
        b = GradBucket(r, ps, alloc)
        assert b.r_numel == 63 and b.mlp_off == 64 and b.grads_end == 64 + 17 and b.scal_off == 84
        assert b.flat.numel() == 84 + GradBucket.N_SCALARS and b.grads.numel() == b.grads_end
        assert b.scalars.data_ptr() == b.flat.data_ptr() + 4 * b.scal_off
        assert [v.shape for v in b.views] == [ps[0].shape, ps[1].shape]      # the frozen parameter has no slot
        b.attach(r)
        assert float(b.flat.abs().sum()) == 0.0                   # attach = zero_grad: also clears external storage
        assert r.grad.data_ptr() == b.flat.data_ptr() and ps[0].grad.data_ptr() == b.flat.data_ptr() + 4 * 64
        (r.sum() * 2 + ps[0].sum() * 3 + ps[1].sum() * 5).backward()
        assert float(b.flat[:63].sum()) == 126.0 and float(b.flat[63]) == 0.0
        assert float(b.flat[64:81].sum()) == 45.0 + 10.0 and float(b.flat[81:].abs().sum()) == 0.0
        r.grad = None
        for q in ps:
            q.grad = None
    # no trainable R: the MLP segment starts at 0
    b = GradBucket(None, ps[:2])
    assert b.r_numel == 0 and b.mlp_off == 0 and b.r_view is None


def test_peer_all_is_a_cuda_feature():
    """peer_all needs CUDA IPC: on CPU tensors the stepper keeps the torch.distributed path (and a single process has
    nothing to exchange); GraphedTrainStep says what it needs under N ranks."""
    args = make_args()
    model, step, _ = build(args)
    from mpvae_b200.train import DataParallelStep
    st = DataParallelStep(model, step.optimizer, None, args, clip_norm=100.0, loss_fn=step.loss_fn, peer_all=True, peer_g_r=True)
    assert st.pbucket is None and st.ring is None
