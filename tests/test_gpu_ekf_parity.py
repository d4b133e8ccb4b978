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
    good = torch.linalg.cond(S) < 1e3
    K = torch.linalg.solve(S, H[sl] @ Pp[sl]).transpose(1, 2)
    err = (Pu[sl] - (Pp[sl] - K @ S @ K.transpose(1, 2))).abs().amax(dim=(1, 2))
    assert good.float().mean().item() > 0.5 and err[good].max().item() <= 1e-9 * Pp.abs().max().item()
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
