"""CPU model of the register-level data flow of armnet_fwd_mma_kernel (armnet_b200/csrc/fused_fwd_mma.cuh).

The kernel chains two warp-level m16n8k8 MMAs through registers: the D fragment of the logits MMA (one thread: two rows
x two adjacent fields per 8-field block) becomes, after the entmax gates, the A fragment of the cross-product MMA under the
k-slot <-> field permutation  t <-> 8j + 2t,  t + 4 <-> 8j + 2t + 1, with B rows permuted the same way.  This test
emulates the PTX fragment layouts of `mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32` for the 32 lanes of a warp in
numpy and checks that the chain, the row permutation (MMA row i <-> neuron 2i / 2(i-8)+1), the leftover-lane handling
(nemb = 8 EK + ER) and the raw-bits TF32 split reproduce  X = M'^T e^T  and  s = w e  of models/armnet.py:33-34,86-87.
It pins the index algebra the header comment states; the kernel itself is parity-tested on the GPU
(tests/test_gpu_parity.py::test_tensor_core_kernel_*)."""
import numpy as np
import pytest

f32 = np.float32


def mma_m16n8k8(d, a, b):
    """d[lane][4] += A(16x8) B(8x8) with the PTX fragment layouts; a[lane][4], b[lane][2] (lane = 4 g + t)."""
    A = np.zeros((16, 8), np.float64)
    B = np.zeros((8, 8), np.float64)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a[lane]
        B[t, g], B[t + 4, g] = b[lane]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        d[lane] += (D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1])


def trunc_tf32(x):
    return (np.asarray(x, f32).view(np.uint32) & np.uint32(0xffffe000)).view(f32)


@pytest.mark.parametrize('F,E', [(39, 10), (40, 10), (33, 10), (39, 16), (36, 16)])
def test_fragment_chain_reproduces_both_products(F, E):
    rng = np.random.default_rng(F * 100 + E)
    NT, EK, ER = (F + 7) // 8, E // 8, E % 8
    assert ER in (0, 2)
    e = rng.standard_normal((F, E))                 # the sample's scaled embedding rows
    Mp = rng.standard_normal((E, 16))               # M'[x][r] for the 16 neurons of one MMA row step
    gate = rng.random((16, F))                      # stands for the entmax gates * values (any function of the logits)

    # thread (g, t) owns neurons r0 = 2g, r1 = 2g + 1 and fields 8j + 2t, 8j + 2t + 1
    X = np.zeros((16, 8 * NT))
    S = np.zeros((16, E))
    c = [np.zeros((32, 4)) for _ in range(NT)]
    for j in range(NT):
        for ks in range(EK):
            a = np.zeros((32, 4))
            b = np.zeros((32, 2))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                r0, r1 = 2 * g, 2 * g + 1
                a[lane] = (Mp[8 * ks + t, r0], Mp[8 * ks + t, r1], Mp[8 * ks + t + 4, r0], Mp[8 * ks + t + 4, r1])
                fB = min(8 * j + g, F - 1)          # padded fields re-read a real row (masked later)
                b[lane] = (e[fB, 8 * ks + t], e[fB, 8 * ks + t + 4])
            mma_m16n8k8(c[j], a, b)
        for lane in range(32):                      # leftover lanes on the FP32 pipe
            g, t = lane >> 2, lane & 3
            for cc in range(2):
                f = min(8 * j + 2 * t + cc, F - 1)
                for h in range(2):
                    for x in range(8 * EK, 8 * EK + ER):
                        c[j][lane][2 * h + cc] += e[f, x] * Mp[x, 2 * g + h]
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for cc in range(2):
                for h in range(2):
                    X[2 * g + h, 8 * j + 2 * t + cc] = c[j][lane][2 * h + cc]
    ref_X = (e @ Mp).T                              # [16, F]
    np.testing.assert_allclose(X[:, :F], ref_X, rtol=1e-12, atol=1e-12)

    # cross product: A fragment = the thread's own (gated) D fragment, k-slot t <-> field 8j+2t, t+4 <-> 8j+2t+1
    acc = [np.zeros((32, 4)) for _ in range(EK)]
    accr = np.zeros((32, 2, max(ER, 1)))
    for j in range(NT):
        a = np.zeros((32, 4))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            w = np.zeros((2, 2))
            for cc in range(2):
                f = 8 * j + 2 * t + cc
                for h in range(2):
                    w[h, cc] = gate[2 * g + h, f] if f < F else 0.0      # padded fields carry zero gates
            a[lane] = (w[0, 0], w[1, 0], w[0, 1], w[1, 1])
            for h in range(2):
                for cc in range(2):
                    f = min(8 * j + 2 * t + cc, F - 1)
                    for xi in range(ER):
                        accr[lane, h, xi] += w[h, cc] * e[f, 8 * EK + xi]
        for n in range(EK):
            b = np.zeros((32, 2))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                b[lane] = (e[min(8 * j + 2 * t, F - 1), 8 * n + g], e[min(8 * j + 2 * t + 1, F - 1), 8 * n + g])
            mma_m16n8k8(acc[n], a, b)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for n in range(EK):
            for h in range(2):
                S[2 * g + h, 8 * n + 2 * t] = acc[n][lane][2 * h]
                S[2 * g + h, 8 * n + 2 * t + 1] = acc[n][lane][2 * h + 1]
    for g in range(8):                              # leftover lanes: quad sum over t
        for h in range(2):
            for xi in range(ER):
                S[2 * g + h, 8 * EK + xi] = sum(accr[4 * g + t, h, xi] for t in range(4))
    np.testing.assert_allclose(S, gate @ e, rtol=1e-12, atol=1e-12)


def test_raw_bits_split_is_fp32_accurate():
    """hi = the fp32 value as the tensor core reads it (top 19 bits), lo = x - trunc(x):  a_lo b_hi + a_hi b_lo + a_hi b_hi
    with every operand truncated to TF32 reproduces an fp32 dot product to ~2^-20 relative."""
    rng = np.random.default_rng(7)
    a = rng.standard_normal((2000, 40)).astype(f32)
    b = rng.standard_normal((2000, 40)).astype(f32)
    a_hi, b_hi = trunc_tf32(a), trunc_tf32(b)
    a_lo, b_lo = trunc_tf32(a - a_hi), trunc_tf32(b - b_hi)
    assert np.all(np.abs(a - a_hi) <= np.abs(a) * 2.0 ** -10)
    got = (a_lo.astype(np.float64) * b_hi + a_hi.astype(np.float64) * b_lo + a_hi.astype(np.float64) * b_hi).sum(1)
    ref = (a.astype(np.float64) * b).sum(1)
    scale = np.abs(a.astype(np.float64) * b).sum(1)
    assert np.max(np.abs(got - ref) / scale) < 4e-6           # worst case 3 x 2^-20; plain TF32 would be ~1e-3
    plain = (a_hi.astype(np.float64) * b_hi).sum(1)
    assert np.max(np.abs(plain - ref) / scale) > 1e-4
