"""Validate the oracle against the reference classes executed live
(AST-sliced from /root/reference).  Only runs in the build container; on the
GPU box the reference is absent and the golden vectors take over."""
import numpy as np
import pytest
import torch

from oracle import ref_slice, vfm_math, vfm_port

pytestmark = pytest.mark.skipif(not ref_slice.available(), reason="/root/reference not present")


def _ids(fs, rows, seed):
    rng = np.random.default_rng(seed)
    offs = np.concatenate(([0], np.cumsum(fs)[:-1]))
    x = np.stack([offs[g] + rng.integers(0, fs[g], rows) for g in range(len(fs))], 1).astype(np.int64)
    y = np.clip(np.round(3.5 + rng.standard_normal(rows)), 1, 5).astype(np.float32)
    return x, y


@pytest.mark.parametrize("output", ["reg", "class"])
@pytest.mark.parametrize("S,link", [(1, "abs"), (3, "abs"), (1, "softplus")])
def test_sampled_port_equals_reference_fp64(output, S, link):
    N, M, d, rows, B = 50, 30, 6, 384, 128
    x, y = _ids([N, M], rows, 1)
    if output == "class":
        y = (y > 3).astype(np.float32)
    tc = np.bincount(x.reshape(-1), minlength=N + M)
    tc[tc == 0] = 1
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        xt, yt, tct = torch.from_numpy(x), torch.from_numpy(y).double(), torch.from_numpy(tc)
        CF = ref_slice.sampled_cf_class(N, M, d, tct, n_var_samples=S, link=link)
        torch.manual_seed(42)
        ref = CF(d, output=output)
        torch.manual_seed(42)
        port = vfm_port.SampledPort(N, M, d, tct, output=output, n_var_samples=S, link=link)
        for (k1, p1), (k2, p2) in zip(ref.named_parameters(), port.named_parameters()):
            assert k1 == k2 and torch.equal(p1, p2)
        o1 = torch.optim.Adam(ref.parameters(), lr=0.05)
        o2 = torch.optim.Adam(port.parameters(), lr=0.05)
        gen = torch.Generator().manual_seed(7)
        for t in range(3):
            xb, yb = xt[t * B:(t + 1) * B], yt[t * B:(t + 1) * B]
            U = len(torch.unique(xb))
            noise = [torch.randn(S, 1, generator=gen), torch.randn(S, U, generator=gen),
                     torch.randn(S, U, d, generator=gen)]
            sd = {k: v.detach().numpy().copy() for k, v in ref.state_dict().items()}
            r1 = ref_slice.sampled_step(ref, o1, xb, yb, rows, noise)
            r2 = vfm_port.sampled_port_step(port, o2, xb, yb, rows, noise)
            P = {"alpha": sd["alpha"], "global_bias_mean": sd["global_bias_mean"],
                 "global_bias_scale": sd["global_bias_scale"],
                 "bias": sd["bias_params.weight"], "entity": sd["entity_params.weight"]}
            r3 = vfm_math.sampled_step(P, xb.numpy(), yb.numpy(), [n.numpy() for n in noise], tc,
                                       rows, [N, M], output=output, link=link)
            assert torch.allclose(r1["loss"], r2["loss"], rtol=1e-12)
            assert torch.allclose(r1["pred"], r2["pred"], rtol=1e-12)
            # reference casts the target to fp32 (vfm-torch.py:359) -> ~1e-8 when run in fp64
            np.testing.assert_allclose(r3["loss"], r1["loss"].numpy()[0], rtol=1e-7)
            for k, kk in [("bias_params.weight", "bias"), ("entity_params.weight", "entity"),
                          ("global_bias_mean",) * 2, ("global_bias_scale",) * 2, ("alpha",) * 2]:
                if r1["grads"][k] is None:
                    assert kk not in r3["grads"]
                    continue
                np.testing.assert_allclose(r3["grads"][kk], r1["grads"][k].numpy(),
                                           rtol=1e-9, atol=1e-9 * np.abs(r3["grads"][kk]).max())
        for (k1, p1), (_, p2) in zip(ref.named_parameters(), port.named_parameters()):
            assert torch.allclose(p1, p2, rtol=1e-12, atol=1e-14), k1
    finally:
        torch.set_default_dtype(old)


@pytest.mark.parametrize("fs", [[50, 30], [3, 30, 50]])
def test_closed_port_equals_reference_fp64(fs):
    G, d, rows, B = len(fs), 6, 384, 128
    x, y = _ids(fs, rows, 2)
    tc = np.bincount(x.reshape(-1), minlength=sum(fs)).astype(np.float64)
    tc[tc == 0] = 1
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        xt, yt, tct = torch.from_numpy(x), torch.from_numpy(y).double(), torch.from_numpy(tc)
        CF = ref_slice.closed_cf_class(fs[0], fs[1])
        torch.manual_seed(42)
        ref = CF(embedding_size=d, n_groups=G, group_sizes=fs, alpha_0=5.)
        torch.manual_seed(42)
        port = vfm_port.ClosedPort(d, fs, alpha_0=5.)
        assert [k for k, _ in ref.named_parameters()] == [k for k, _ in port.named_parameters()]
        gen = torch.Generator().manual_seed(1)
        bump = 0.3 * torch.randn(sum(fs), d, generator=gen)
        with torch.no_grad():
            for mdl in (ref, port):
                mdl.entity_params[:, :d] += bump
        o1 = torch.optim.Adam(ref.parameters(), lr=0.1)
        o2 = torch.optim.Adam(port.parameters(), lr=0.1)
        for t in range(3):
            xb, yb = xt[t * B:(t + 1) * B], yt[t * B:(t + 1) * B]
            sd = {k: v.detach().numpy().copy() for k, v in ref.state_dict().items()}
            r1 = ref_slice.closed_step(ref, o1, xb, yb, rows, tct, fs)
            r2 = vfm_port.closed_port_step(port, o2, xb, yb, rows, tct)
            assert torch.allclose(r1["loss"], r2["loss"], rtol=1e-12)
            assert torch.allclose(r1["pred"], r2["pred"], rtol=1e-12)
            import golden_util as gu
            r3 = vfm_math.closed_step(gu.closed_math_params(sd, G), xb.numpy(), yb.numpy(), tc, rows, fs)
            np.testing.assert_allclose(r3["loss"], r1["loss"].numpy(), rtol=1e-10)
            np.testing.assert_allclose(r3["grads"]["entity"], r1["grads"]["entity_params"].numpy(),
                                       rtol=1e-8, atol=1e-9 * np.abs(r3["grads"]["entity"]).max())
            np.testing.assert_allclose(r3["grads"]["bias"], r1["grads"]["bias_params"].numpy(),
                                       rtol=1e-8, atol=1e-9 * np.abs(r3["grads"]["bias"]).max())
        for (k1, p1), (_, p2) in zip(ref.named_parameters(), port.named_parameters()):
            assert torch.allclose(p1, p2, rtol=1e-12, atol=1e-14), k1
    finally:
        torch.set_default_dtype(old)


def test_golden_files_are_current():
    """The committed goldens were produced by this reference checkout: regenerate
    one small case in memory and compare."""
    import golden_util as gu
    meta, g = gu.load("sampled_reg_d5")
    tc = torch.from_numpy(g["train_counts"])
    CF = ref_slice.sampled_cf_class(meta["N"], meta["M"], meta["d"], tc)
    torch.manual_seed(42)
    ref = CF(meta["d"], output=meta["output"])
    for k, v in gu.state(g, "init").items():
        assert np.array_equal(ref.state_dict()[k].numpy(), v), k
    opt = torch.optim.Adam(ref.parameters(), lr=meta["lr"])
    x, y = gu.batch_of(meta, g, 0)
    noise = [torch.from_numpy(g[f"step0.noise{i}"]) for i in range(3)]
    out = ref_slice.sampled_step(ref, opt, torch.from_numpy(x), torch.from_numpy(y), meta["n_train"], noise)
    np.testing.assert_allclose(out["loss"].numpy(), g["step0.loss"], rtol=1e-6)
    np.testing.assert_allclose(out["pred"].numpy(), g["step0.pred"], rtol=1e-6)


def test_data_loading_equals_reference_prepare_py(tmp_path, monkeypatch):
    """vae_b200.data.load_data == the reference's prepare_data + load_data (prepare.py:10-62), executed
    live on a raw ratings file: ids re-indexed to 0..N-1 / 0..M-1, items shifted by N, fold files
    honoured, 'outcome' = rating >= 4, libFM export lines."""
    import importlib.util
    import os
    import shutil

    import pandas as pd

    from vae_b200 import data
    rng = np.random.default_rng(3)
    n = 400
    raw = pd.DataFrame({"user": rng.choice([3, 7, 8, 15, 21, 22, 40, 41, 97], n),
                        "item": rng.choice([100, 101, 105, 230, 231, 999], n),
                        "rating": rng.integers(1, 6, n)})
    perm = rng.permutation(n)
    ours_dir = tmp_path / "ours"
    ref_dir = tmp_path / "data" / "toybinary"
    for d in (ours_dir, ref_dir):
        os.makedirs(d)
        raw.assign(outcome=(raw["rating"] >= 4).astype(int)).to_csv(d / "data.csv", index=False) if d is ours_dir \
            else raw.to_csv(d / "data.csv", index=False)
        pd.DataFrame({"index": np.sort(perm[:320])}).to_csv(d / "trainval.csv", index=False)
        pd.DataFrame({"index": np.sort(perm[320:])}).to_csv(d / "test.csv", index=False)
    spec = importlib.util.spec_from_file_location("ref_prepare", os.path.join(ref_slice.REFERENCE_DIR, "prepare.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    monkeypatch.chdir(tmp_path)                              # the reference reads ./data/<name>/
    ref.prepare_data("toybinary", True)                      # rewrites data.csv, writes the libFM files
    for output in ("class", "reg"):
        N, M, X_train, X_test, y_train, y_test, folds = ref.load_data("toybinary", output)
        got = data.load_data(str(ours_dir), output)
        assert (got.n_users, got.n_items) == (N, M)
        assert np.array_equal(got.x_train, X_train) and np.array_equal(got.x_test, X_test)
        assert np.array_equal(got.y_train, y_train.astype(np.float32)) and np.array_equal(got.y_test, y_test.astype(np.float32))
        assert np.array_equal(got.folds["trainval"], folds["trainval"]) and np.array_equal(got.folds["test"], folds["test"])
    got = data.load_data(str(ours_dir), "class")
    data.write_libfm(str(tmp_path / "ours.libfm"), got.x_test, got.y_test.astype(int))
    assert (tmp_path / "ours.libfm").read_text() == (ref_dir / "toybinary.test_libfm").read_text()
    shutil.rmtree(tmp_path / "data")
