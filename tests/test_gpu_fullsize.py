"""Oracle parity of the CUDA path at the FULL sizes of BASELINE.json's configs 2-5 (config 1 is already full size in
tests/test_gpu_parity.py::test_detergent_config1_five_objectives and tests/test_gpu_strategy.py).  Each test scores a
sample of the config's raw-sample q-batches (256-2048, as many as the CPU oracle finishes in seconds) on the device
through the C ABI and on the CPU oracle and compares with the tolerance bench.py's cpu_baseline leg uses:
acquisition values 1e-9 of the largest value (1e-8 absolute for log-space values), posterior mean 1e-9 relative.

    python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q
"""
import pytest
import torch

from everest_b200 import configs as Cf
from tests import problems as P

pytestmark = pytest.mark.gpu
DT = torch.float64


def rel_to_max(a, b):
    a, b = a.detach().cpu().to(DT), b.detach().cpu().to(DT)
    return float((a - b).abs().max() / b.abs().max())


def test_config3_zdt1_30d_qnehvi_full_size():
    """N=2000, d=30, 2 outputs, q=4, S=512: baseline pruning (2048 joint samples over all 2000 points), per-sample fronts and
    cells, and 2048 of the 16384 raw-sample q-batches -- enough rows (8192 x 2000 >= 2^22) for the INT8 digit-plane GEMM with
    its per-row guard to be the kernel under test."""
    p = Cf.zdt1_qnehvi()
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp)                         # full oracle set-up, own pruning
    acq_d = Cf.build_acqf(p, st)
    assert acq_d.prune_idx.cpu().tolist() == acq_o.prune_idx.tolist()
    assert acq_d.nb == acq_o.nb and acq_d.nb > 0
    lo_d, up_d, nc_d = acq_d.cell_bounds()
    assert nc_d.tolist() == acq_o.n_cells.tolist()
    for s_ in range(p["S"]):
        c = int(nc_d[s_])
        # identical cell lists whenever the baseline samples agree bit for bit; the samples themselves agree to rounding
        assert float((lo_d[s_, :c] - acq_o.cell_lower[s_, :c]).abs().max()) < 1e-9, s_
        assert float((up_d[s_, :c] - acq_o.cell_upper[s_, :c]).abs().max()) < 1e-9 or bool(torch.isinf(up_d[s_, :c]).any()), s_
        fin = torch.isfinite(acq_o.cell_upper[s_, :c])
        assert torch.equal(torch.isfinite(up_d[s_, :c]), fin)
        assert float((up_d[s_, :c][fin] - acq_o.cell_upper[s_, :c][fin]).abs().max()) < 1e-9, s_
    X = Cf.candidates(p)[:2048]
    v_o, parts = acq_o.forward(X, return_parts=True)
    v_d = acq_d(X.to(st.device)).cpu()
    chk = st.debug_get("ozaki_check", capacity=16).tolist()
    assert chk[0] == 1.0 and chk[2] == 2048 and chk[1] <= 0.01 * 2048   # INT8 path ran; the guard redid a handful at most
    assert rel_to_max(v_d, v_o) < 1e-9
    mu_d = st.debug_get("mu", capacity=2048 * 4 * 2).view(2048, 4, 2).cpu()
    assert float(((mu_d - parts["mu"]).abs() / (parts["mu"].abs() + 1e-6)).max()) < 1e-9
    # the whole 16384-q-batch screen from HOST rows (15.7 MB: the piece mode of bo_acqf_forward_host -- point preparation and
    # K(X*,X) launched piece by piece as the copies land, everything else once over the batch) against the device-resident
    # call: the same arithmetic per element, so the same bits
    X_all = Cf.candidates(p)
    v_host = torch.as_tensor(acq_d.forward_host(X_all.contiguous().numpy()))
    v_dev = acq_d(X_all.to(st.device)).cpu()
    assert torch.equal(v_host, v_dev)
    assert rel_to_max(v_dev[:2048], v_d) < 1e-12
    # the same q-batches through the FP64 kernel
    acq_d.set_option("ozaki", 0)
    assert rel_to_max(acq_d(X.to(st.device)).cpu(), v_o) < 1e-9
    # marginal posterior (BotorchStrategy._predict) on the candidate points
    Xq = X[:256].reshape(-1, p["d"])
    mean_o, cov_o = gp.posterior(Xq)
    mean_d, var_d = st.posterior(Xq)
    var_o = torch.diagonal(cov_o, dim1=-2, dim2=-1).transpose(0, 1)
    assert float(((mean_d.cpu() - mean_o).abs() / (mean_o.abs() + 1e-6)).max()) < 1e-9
    assert float(((var_d.cpu() - var_o).abs() / var_o).max()) < 1e-9


def test_config2_himmelblau_qlogei_full_size():
    """N=500, Scale(Matern-5/2 ARD), qLogEI, q=1, S=512: 2048 of the 4096 raw samples."""
    p = Cf.himmelblau_qlogei()
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp)
    acq_d = Cf.build_acqf(p, st)
    assert abs(acq_d.best_f - acq_o.best_f) <= 1e-9 * max(1.0, abs(acq_o.best_f))
    X = Cf.candidates(p)[:2048]
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device)).cpu()
    assert bool(torch.isfinite(v_o).all())
    assert float((v_d - v_o).abs().max()) < 1e-8 * max(1.0, float(v_o.abs().max()))
    mean_o, cov_o = gp.posterior(X[:512, 0])
    mean_d, var_d = st.posterior(X[:512, 0])
    assert float(((mean_d.cpu() - mean_o).abs() / (mean_o.abs() + 1e-6)).max()) < 1e-9
    var_o = torch.diagonal(cov_o, dim1=-2, dim2=-1).transpose(0, 1)
    assert float(((var_d.cpu() - var_o).abs() / var_o).max()) < 1e-9


def test_config4_dtlz2_4obj_q8_full_size():
    """N=1000, d=6, 4 objectives, q=8 (255 subsets), S=512, thousands of cells per MC sample: 16 q-batches on the oracle
    (S x C x 255 subset terms each), with the device's cell lists injected into the oracle (its pure-Python decomposition
    of 512 fronts takes minutes; the lists are pinned bit for bit at small sizes) and a hypervolume cross-check of the
    cells that needs no oracle: for every tenth MC sample, sum of cell volumes == box(ref, ideal) - dominated HV."""
    from everest_b200 import multiobjective as MO
    from everest_b200.objectives import MaximizeObjective, MultiObjective

    p = Cf.dtlz2_qnehvi()
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_d = Cf.build_acqf(p, st)
    assert acq_d.max_cells > 100
    acq_o = P.oracle_qnehvi_on_device_baseline(p, gp, acq_d, inject_cells=True)
    # baseline samples and objectives of the two paths agree before any cell is used
    obj_d = st.debug_get("obj_b", capacity=p["S"] * acq_d.nb * 4).view(p["S"], acq_d.nb, 4).cpu()
    assert float((obj_d - acq_o.obj_b).abs().max()) < 1e-9 * float(acq_o.obj_b.abs().max())
    # the cells of a sample partition the non-dominated part of [ref, ideal]
    lo, up, nc = acq_d.cell_bounds()
    ref = torch.tensor(p["ref_point"], dtype=DT)
    mobj = MultiObjective([MaximizeObjective(i) for i in range(4)])
    for s_ in range(0, p["S"], 50):
        Y = obj_d[s_]
        ideal = torch.maximum(Y.max(dim=0).values, ref)
        vol = ((torch.minimum(up[s_, : nc[s_]], ideal) - lo[s_, : nc[s_]]).clamp_min(0.0)).prod(dim=-1).sum()
        hv = MO.compute_hypervolume(mobj, Y.numpy(), p["ref_point"])
        assert abs(float(vol) - (float((ideal - ref).prod()) - hv)) < 1e-9 * float((ideal - ref).prod()), s_
    X = Cf.candidates(p)[:16]
    v_o = acq_o.forward(X)
    v_d = acq_d(X.to(st.device)).cpu()
    assert float(v_o.abs().max()) > 0
    assert rel_to_max(v_d, v_o) < 1e-9


def test_config5_mixed_tanimoto_hamming_full_size():
    """N=5000, 2048-bit fingerprints (bit-packed popcount Tanimoto), two one-hot categoricals (Hamming), two continuous
    columns, composite (sum + product) kernel tree, qLogEI over a discrete choice set: 1024 choices."""
    p = Cf.mixed_tanimoto_qlogei(n_choices=4096)
    gp = P.oracle_gp(p)
    st = Cf.build_state(p)
    acq_o = P.oracle_acqf(p, gp)
    acq_d = Cf.build_acqf(p, st)
    assert abs(acq_d.best_f - acq_o.best_f) <= 1e-9 * max(1.0, abs(acq_o.best_f))
    X = Cf.candidates(p)
    v_d = acq_d(X.to(st.device)).cpu()                    # 4096 x 5000 >= 2^22: the INT8 GEMM with its guard
    v_o = acq_o.forward(X[:1024])
    assert float((v_d[:1024] - v_o).abs().max()) < 1e-8 * max(1.0, float(v_o.abs().max()))
    mean_o, cov_o = gp.posterior(X[:256, 0])
    mean_d, var_d = st.posterior(X[:256, 0])
    assert float(((mean_d.cpu() - mean_o).abs() / (mean_o.abs() + 1e-6)).max()) < 1e-9
    var_o = torch.diagonal(cov_o, dim1=-2, dim2=-1).transpose(0, 1)
    assert float(((var_d.cpu() - var_o).abs() / var_o).max()) < 1e-9
    # arg-max of the discrete branch (optimize_acqf_discrete, botorch.py:461) agrees on the shared sample
    assert int(torch.argmax(v_d[:1024])) == int(torch.argmax(v_o))


def test_packed_wire_format_of_the_host_entry_point():
    """Config-5 layout (2 continuous | 2048 fingerprint bits | 4 + 6 one-hot): the host entry point ships the fingerprint
    block as bits (bo_acqf_forward_host packs while staging; bo_pack_layout / bo_pack_rows_host / bo_acqf_forward_host_packed
    for callers that keep the choice set packed).  The device restores exactly the same 0 / 1 doubles, so the values agree with
    the device-pointer call on the float64 rows to the rounding that the chunking of the host pipeline introduces (the
    Gram partial sums are grouped per chunk); a non-binary fingerprint is refused."""
    import numpy as np

    p = Cf.mixed_tanimoto_qlogei(N=300, S=64, n_choices=6000)
    st = Cf.build_state(p)
    acq = Cf.build_acqf(p, st)
    X = Cf.candidates(p)                              # [6000, 1, 2060]: 99 MB of float64, > 8 MiB: the chunked pipeline
    v_dev = acq(X.to(st.device)).cpu().numpy()
    dc, bc = acq.pack_layout()
    assert bc.tolist() == list(range(2, 2050)) and dc.tolist() == [0, 1] + list(range(2050, 2060))
    v_host = acq.forward_host(X.numpy())             # packs inside
    tol = dict(rtol=1e-9, atol=1e-11 * float(np.abs(v_dev).max()))
    assert np.allclose(v_host, v_dev, **tol)
    dense, bits = acq.pack_rows(X.numpy())
    assert dense.shape == (6000, 12) and bits.shape == (6000, 32) and bits.dtype == np.uint64
    # the library's packer against numpy's: bit k of a row = column bit_cols[k], little-endian inside each 64-bit word
    ref_bits = np.ascontiguousarray(np.packbits(X.numpy().reshape(6000, -1)[:, bc].astype(np.uint8), axis=1,
                                                bitorder="little")).view(np.uint64)
    assert np.array_equal(bits, ref_bits)
    v_packed = acq.forward_host_packed(dense, bits, q=1)
    assert np.allclose(v_packed, v_dev, **tol)
    v_small = acq.forward_host_packed(dense[:7], bits[:7], q=1)      # single chunk
    assert np.allclose(v_small, v_dev[:7], **tol)
    bad = X.numpy().copy()
    bad[5, 0, 100] = 0.5
    with pytest.raises(ValueError):
        acq.forward_host(bad)
    with pytest.raises(ValueError):
        acq.pack_rows(bad)
