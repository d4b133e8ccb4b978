"""GPU parity tests for the EKF covariance algebra: CUDA engine (through the C ABI) vs the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_F64 = 1e-6  # north_star tolerance for EKF states, fp64 (observed agreement is ~1e-13)


@pytest.fixture(scope="module")
def sfb():
    import smooth_feedback_b200 as s

    return s


def dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


def cm(a):
    return np.ascontiguousarray(np.swapaxes(a, -1, -2))


def relmax(a, ref):
    return np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("d", [1, 2, 3, 6, 9, 10])
@pytest.mark.parametrize("stepper,dt", [("euler", None), ("euler", 0.03), ("rk4", 0.01)])
def test_predict_parity(sfb, oracle, d, stepper, dt):
    from smooth_feedback_b200.generators import random_ekf_numpy

    B = 1000 + d  # ragged tail tile
    P, A, Q, _, _, _ = random_ekf_numpy(B, d, 3, seed=d)
    out = sfb.ekf_predict_batch(dev(cm(P)), dev(cm(A)), dev(cm(Q)), 0.1, dt, stepper)
    got = np.swapaxes(out.cpu().numpy(), -1, -2)
    ref = oracle.ekf_predict_batch(P, A, Q, 0.1, dt, stepper, nthreads=8)
    assert relmax(got, ref) <= 1e-12
    assert np.array_equal(got, np.swapaxes(got, -1, -2))  # selfadjointView<Upper> output is exactly symmetric


@pytest.mark.parametrize("d,ny", [(3, 3), (10, 3), (3, 10), (6, 3), (6, 6), (1, 1), (2, 5)])
def test_update_parity(sfb, oracle, d, ny):
    from smooth_feedback_b200.generators import random_ekf_numpy

    B = 777
    P, _, _, H, R, innov = random_ekf_numpy(B, d, ny, seed=d * 31 + ny)
    delta, Pn = sfb.ekf_update_batch(dev(cm(P)), dev(cm(H)), dev(cm(R)), dev(innov))
    got_d = delta.cpu().numpy(); got_P = np.swapaxes(Pn.cpu().numpy(), -1, -2)
    ref_d, ref_P = oracle.ekf_update_batch(P, H, R, innov, nthreads=8)
    assert relmax(got_d, ref_d) <= 1e-9 and relmax(got_P, ref_P) <= 1e-9
    # textbook Kalman update at the reference test's tolerance (tests/test_ekf.cpp:96-98)
    S = H @ P @ np.swapaxes(H, 1, 2) + R
    K = P @ np.swapaxes(H, 1, 2) @ np.linalg.inv(S)
    assert relmax(got_d, np.einsum("bij,bj->bi", K, innov)) <= REL_F64
    assert relmax(got_P, (np.eye(d) - K @ H) @ P) <= REL_F64


def test_full_size_properties(sfb, oracle):
    """BASELINE.json configs[3] shape (d=6, ny=3, batch 2^20, fp64): size-independent properties."""
    import torch

    B, d, ny = 1 << 20, 6, 3
    g = torch.Generator(device="cuda").manual_seed(5)
    M = torch.rand(B, d, d, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    P = M @ M.transpose(1, 2) + 0.1 * torch.eye(d, device="cuda", dtype=torch.float64)
    P = 0.5 * (P + P.transpose(1, 2))
    A = torch.randn(B, d, d, generator=g, device="cuda", dtype=torch.float64)
    Q = (0.01 * torch.eye(d, device="cuda", dtype=torch.float64)).expand(B, d, d).contiguous()
    tau = 0.1
    Pp = sfb.ekf_predict_batch(P, A.transpose(1, 2).contiguous(), Q, tau)
    ref = P + tau * (A @ P + P @ A.transpose(1, 2) + Q)
    assert (Pp - ref).abs().max().item() <= 1e-12 * ref.abs().max().item()
    assert torch.equal(Pp, Pp.transpose(1, 2))
    H = torch.randn(B, ny, d, generator=g, device="cuda", dtype=torch.float64)
    R = (0.01 * torch.eye(ny, device="cuda", dtype=torch.float64)).expand(B, ny, ny).contiguous()
    innov = torch.randn(B, ny, generator=g, device="cuda", dtype=torch.float64)
    delta, Pu = sfb.ekf_update_batch(Pp, H.transpose(1, 2).contiguous(), R, innov)
    torch.cuda.synchronize()
    assert torch.equal(Pu, Pu.transpose(1, 2))                                   # selfadjointView<Upper>: exactly symmetric
    # NB: (I - K H) P is the reference's update (ekf.hpp:138), not the Joseph form: it does not guarantee a positive
    # definite result when S is ill conditioned (cond(S) reaches 1e6 in 2^20 draws), and neither does the reference.
    # Size-independent identity on a well-conditioned strided sample: Pu = Pp - K S K^T, PD, trace never grows.
    sl = slice(0, B, 64)
    S = H[sl] @ Pp[sl] @ H[sl].transpose(1, 2) + R[sl]
    # the Euler step P + tau (A P + P A^T + Q) drops tau^2 A P A^T, so Pp itself is not always positive definite;
    # positive definiteness of Pu = (Pp^-1 + H^T R^-1 H)^-1 is only a property of instances whose Pp is
    good = (torch.linalg.cond(S) < 1e3) & (torch.linalg.eigvalsh(Pp[sl]).amin(dim=1) > 1e-3)
    K = torch.linalg.solve(S, H[sl] @ Pp[sl]).transpose(1, 2)
    err = (Pu[sl] - (Pp[sl] - K @ S @ K.transpose(1, 2))).abs().amax(dim=(1, 2))
    assert good.float().mean().item() > 0.2 and err[good].max().item() <= 1e-9 * Pp.abs().max().item()
    assert (torch.linalg.cholesky_ex(Pu[sl][good]).info == 0).all()
    tr = lambda M: M.diagonal(dim1=1, dim2=2).sum(1)
    assert (tr(Pu[sl])[good] <= tr(Pp[sl])[good]).all()
    # parity with the oracle on a strided sample
    pick = torch.arange(0, B, 257, device="cuda")
    c = lambda t: t[pick].cpu().numpy()
    oPp = oracle.ekf_predict_batch(c(P), c(A), c(Q), tau, nthreads=8)
    od, oPu = oracle.ekf_update_batch(oPp, c(H), c(R), c(innov), nthreads=8)
    assert relmax(c(Pp), oPp) <= 1e-12
    assert relmax(c(Pu), oPu) <= REL_F64 and relmax(c(delta), od) <= REL_F64


@pytest.mark.parametrize("d,ny", [(6, 3), (6, 6), (3, 3), (4, 2), (6, 1), (6, 2), (3, 1), (2, 2), (5, 2)])
@pytest.mark.parametrize("B", [1, 63, 64, 65, 1000, 12345])
def test_fused_step_parity(sfb, oracle, d, ny, B):
    """sfb_ekf_step_batch_f64 (TMA-staged fused predict+update; (5,2) takes the generic kernels) == oracle predict->update,
    including ragged last tiles (B not a multiple of the 64-instance tile)."""
    from smooth_feedback_b200.generators import random_ekf_numpy

    P, A, Q, H, R, innov = random_ekf_numpy(B, d, ny, seed=11 + B)
    Q = Q + 0.003 * np.triu(np.ones((d, d)))[None]      # non-trivial upper triangle, lower ignored by symU
    R = R + 0.002 * np.triu(np.ones((ny, ny)))[None]
    for dt in (None, 0.03):
        delta, Pu = sfb.ekf_step_batch(dev(cm(P)), dev(cm(A)), dev(cm(Q)), 0.1, dev(cm(H)), dev(cm(R)), dev(innov), dt=dt)
        oPp = oracle.ekf_predict_batch(P, A, Q, 0.1, dt=dt)
        od, oPu = oracle.ekf_update_batch(oPp, H, R, innov)
        got_P = np.swapaxes(Pu.cpu().numpy(), 1, 2)
        # per-instance relative errors: the euler-predicted covariance can be indefinite, so S = H P H^T + R is
        # occasionally near-singular (most often when ny = d) and amplifies last-bit differences (FMA contraction,
        # 1/D formed once) by cond(S).  North-star tolerance (1e-6) on every instance, 1e-9 on 99 % of them.
        eP = np.abs(got_P - oPu).max(axis=(1, 2)) / np.abs(oPu).max(axis=(1, 2))
        ed = np.abs(delta.cpu().numpy() - od).max(axis=1) / np.abs(od).max(axis=1)
        Pp_s = np.triu(oPp) + np.swapaxes(np.triu(oPp, 1), 1, 2)
        condS = np.linalg.cond(H @ Pp_s @ np.swapaxes(H, 1, 2) + np.triu(R) + np.swapaxes(np.triu(R, 1), 1, 2))
        ok = condS < 1e6
        assert eP[ok].max() <= REL_F64 and ed[ok].max() <= REL_F64, (eP.max(), ed.max())
        assert np.quantile(eP, 0.99) <= 1e-9 and np.quantile(ed, 0.99) <= 1e-9
        assert ok.mean() > 0.98
        assert np.array_equal(got_P, np.swapaxes(got_P, 1, 2))


def test_fused_step_in_place_and_equals_two_calls(sfb):
    """out_P may alias P; fused == predict followed by update to the last bit is not required, 1e-12 is."""
    from smooth_feedback_b200.generators import random_ekf_numpy

    B, d, ny = 4099, 6, 3
    P, A, Q, H, R, innov = random_ekf_numpy(B, d, ny, seed=3)
    tP, tA, tQ, tH, tR, ti = dev(cm(P)), dev(cm(A)), dev(cm(Q)), dev(cm(H)), dev(cm(R)), dev(innov)
    Pp = sfb.ekf_predict_batch(tP, tA, tQ, 0.1)
    d2, P2 = sfb.ekf_update_batch(Pp, tH, tR, ti)
    Pin = tP.clone()
    d1, P1 = sfb.ekf_step_batch(Pin, tA, tQ, 0.1, tH, tR, ti, out_P=Pin)
    assert P1.data_ptr() == Pin.data_ptr()
    assert relmax(P1.cpu().numpy(), P2.cpu().numpy()) <= 1e-12 and relmax(d1.cpu().numpy(), d2.cpu().numpy()) <= 1e-12


def test_host_pointer_path(sfb, oracle):
    """All-host pointers through the C ABI (staged round trip) == the device path == the oracle."""
    from smooth_feedback_b200.generators import random_ekf_numpy

    B, d, ny = 777, 6, 3
    P, A, Q, H, R, innov = random_ekf_numpy(B, d, ny, seed=21)
    dh, Ph = sfb.ekf_step_batch_host(cm(P), cm(A), cm(Q), 0.1, cm(H), cm(R), innov)
    dd, Pd = sfb.ekf_step_batch(dev(cm(P)), dev(cm(A)), dev(cm(Q)), 0.1, dev(cm(H)), dev(cm(R)), dev(innov))
    assert np.array_equal(dh, dd.cpu().numpy()) and np.array_equal(Ph, Pd.cpu().numpy())
    oPp = oracle.ekf_predict_batch(P, A, Q, 0.1)
    od, oPu = oracle.ekf_update_batch(oPp, H, R, innov)
    assert relmax(np.swapaxes(Ph, 1, 2), oPu) <= 1e-9 and relmax(dh, od) <= 1e-9
