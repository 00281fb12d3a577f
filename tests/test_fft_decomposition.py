"""The register-resident frontend kernel (csrc/frontend.cu logmel_reg_kernel) computes the 512-point real FFT of a frame as a 256-point
complex FFT factored 256 = 8 x 8 x 4 over the 32 lanes of a warp (8 points per lane), with two transposes through shared memory and
the even/odd split done by pairing register i of lane l with register 7 - i of lane (32 - l) mod 32.  This CPU test executes exactly
that data flow in numpy -- lane / register indices, twiddle exponents, transpose addresses (pitches 34 and 40 float2) and their
bank-conflict freedom, the shuffle pairing -- and checks it against numpy's FFT.  It pins the ALGORITHM (index algebra); the CUDA
code is checked on the GPU against the oracle (tests/test_gpu_frontend.py) and against the shared-memory Stockham kernel
(tests/test_gpu_ab_switches.py).  Reference semantics: rust/features/src/lib.rs:66-120 (512-point real FFT of the windowed frame)."""
import numpy as np

T1_PITCH, T2_PITCH = 34, 40


def W(n, e):
    return np.exp(-2j * np.pi * e / n)


def dft4(x0, x1, x2, x3):
    s0, d0, s1, d1 = x0 + x2, x0 - x2, x1 + x3, x1 - x3
    return [s0 + s1, d0 - 1j * d1, s0 - s1, d0 + 1j * d1]


def dft8(a):
    e = dft4(a[0], a[2], a[4], a[6])
    o = dft4(a[1], a[3], a[5], a[7])
    r = np.sqrt(0.5)
    w = [1, r * (1 - 1j), -1j, r * (-1 - 1j)]
    out = [0] * 8
    for k in range(4):
        t = w[k] * o[k]
        out[k], out[k + 4] = e[k] + t, e[k] - t
    return out


def conflict_free(addrs):
    """a 64-bit shared-memory access is served per half-warp: 16 lanes must hit 16 distinct 8-byte bank pairs"""
    for h in range(2):
        assert len({a % 16 for a in addrs[16 * h:16 * h + 16]}) == 16, addrs


def warp_fft256(x):
    reg = np.array([[x[l + 32 * i] for i in range(8)] for l in range(32)])
    for l in range(32):                                       # pass 1: radix-8 over the register index, twiddle W_256^(lane k1)
        y = dft8(list(reg[l]))
        reg[l] = [y[k] * W(256, l * k) for k in range(8)]
    buf = np.zeros(8 * T2_PITCH, complex)
    for k in range(8):
        ad = [T1_PITCH * k + l for l in range(32)]
        conflict_free(ad)
        buf[ad] = reg[:, k]
    r2 = np.zeros((32, 8), complex)
    for j in range(8):                                        # transpose 1: lane k1 + 8 n00 gathers n01 = 0..7
        ad = [T1_PITCH * (l & 7) + (l >> 3) + 4 * j for l in range(32)]
        conflict_free(ad)
        r2[:, j] = buf[ad]
    for l in range(32):                                       # pass 2: radix-8, twiddle W_32^(n00 k20)
        y = dft8(list(r2[l]))
        r2[l] = [y[k] * W(32, (l >> 3) * k) for k in range(8)]
    buf[:] = 0
    for k in range(8):
        ad = [T2_PITCH * k + l for l in range(32)]
        conflict_free(ad)
        buf[ad] = r2[:, k]
    z = np.zeros((32, 8), complex)
    for j in range(2):                                        # transpose 2 + pass 3: two radix-4 per lane
        src = []
        for n00 in range(4):
            ad = [T2_PITCH * ((l >> 3) + 4 * j) + (l & 7) + 8 * n00 for l in range(32)]
            conflict_free(ad)
            src.append(buf[ad])
        for l in range(32):
            out = dft4(src[0][l], src[1][l], src[2][l], src[3][l])
            for k21 in range(4):
                z[l, j + 2 * k21] = out[k21]
    return z                                                  # z[l, i] = Z[l + 32 i]


def test_three_pass_register_fft_equals_numpy():
    rng = np.random.default_rng(0)
    for _ in range(4):
        x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
        z = warp_fft256(x)
        got = np.array([z[k % 32, k // 32] for k in range(256)])
        assert np.abs(got - np.fft.fft(x)).max() < 1e-12


def test_even_odd_split_by_lane_pairing_equals_rfft():
    rng = np.random.default_rng(1)
    frame = np.zeros(512)
    frame[:400] = rng.standard_normal(400)                    # 400 windowed samples, tail zero-padded (lib.rs:66-120)
    z = warp_fft256(frame[0::2] + 1j * frame[1::2])
    X = np.zeros(257, complex)
    for l in range(32):
        for i in range(8):
            zk = z[l, i]
            zn = z[(32 - l) & 31, 7 - i] if l else z[0, (8 - i) & 7]
            e = 0.5 * (zk + np.conj(zn))
            o = -0.5j * (zk - np.conj(zn))
            X[l + 32 * i] = e + o * W(512, l) * W(16, i)
    X[256] = z[0, 0].real - z[0, 0].imag
    assert np.abs(X - np.fft.rfft(frame)).max() < 1e-12


def test_points_beyond_the_window_are_the_lanes_the_kernel_skips():
    """packed point n = lane + 32 i holds samples 2n, 2n+1: non-zero only for n < 200, i.e. i <= 5 everywhere, i == 6 for lanes < 8"""
    nz = {(l, i) for l in range(32) for i in range(8) if 2 * (l + 32 * i) < 400}
    assert nz == {(l, i) for l in range(32) for i in range(6)} | {(l, 6) for l in range(8)}
