"""Lane-accurate CPU emulation of the index math of narrow_out_mma_kernel / narrow_in_mma_kernel
(csrc/conv_narrow.cu): every lane builds its mma.m16n8k16 fragments with the kernel's formulas,
the MMA itself is emulated from the documented fragment layout, and the result is compared with
torch conv2d.  Checks the GEMM-K / GEMM-N permutations, tap handling and the store mapping without
a GPU (values stay fp32: the 16-bit hi/lo split is not modelled).  Run: python scripts/emulate_narrow_mma.py
"""
import numpy as np
import torch
import torch.nn.functional as F


def mma_16816(a_frag, b_frag, d_frag):
    """a_frag [32][4][2], b_frag [32][2][2], d_frag [32][4] -> d_frag += A @ B by fragment layout."""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for e in range(2):
            A[g, 2 * t + e] = a_frag[lane][0][e]
            A[g + 8, 2 * t + e] = a_frag[lane][1][e]
            A[g, 2 * t + 8 + e] = a_frag[lane][2][e]
            A[g + 8, 2 * t + 8 + e] = a_frag[lane][3][e]
            B[2 * t + e, g] = b_frag[lane][0][e]
            B[2 * t + 8 + e, g] = b_frag[lane][1][e]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        d_frag[lane][0] += D[g, 2 * t]
        d_frag[lane][1] += D[g, 2 * t + 1]
        d_frag[lane][2] += D[g + 8, 2 * t]
        d_frag[lane][3] += D[g + 8, 2 * t + 1]


def narrow_out_emul(x_nhwc, wt, CN, Ho, Wo, taps):
    """x_nhwc [N][Hi][Wi][64], wt [tap][64][CN]; taps = list of (dh, dw)."""
    N, Hi, Wi, _ = x_nhwc.shape
    NT = 1 if CN <= 8 else 2
    n_taps = len(taps)
    out = np.zeros((N, CN, Ho, Wo))
    # B fragments [ks][nt][lane] -> (b0 pair, b1 pair)
    sb = {}
    for ks in range(16):
        for nt in range(NT):
            for l in range(32):
                g, t = l >> 2, l & 3
                tap, s = ks >> 2, ks & 3
                base = (s >> 1) * 32 + 8 * t + (s & 1) * 4
                co = nt * 8 + g
                w = [wt[tap, base + e, co] if (co < CN and tap < n_taps) else 0.0 for e in range(4)]
                sb[(ks, nt, l)] = ((w[0], w[1]), (w[2], w[3]))
    npix = N * Ho * Wo
    tiles = (npix + 15) // 16
    for tile in range(tiles):
        lanes = []
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            st = {}
            st["pv"], st["pn"], st["pi"], st["pj"] = [], [], [], []
            for r in range(2):
                p = tile * 16 + g + 8 * r
                pv = p < npix
                pp = p if pv else 0
                q = pp // Wo
                st["pv"].append(pv)
                st["pj"].append(pp - q * Wo)
                st["pn"].append(q // Ho)
                st["pi"].append(q - (q // Ho) * Ho)
            v = {}
            for tap in range(4):
                for r in range(2):
                    dh, dw = taps[tap] if tap < n_taps else (0, 0)
                    h, w = st["pi"][r] + dh, st["pj"][r] + dw
                    ok = st["pv"][r] and tap < n_taps and 0 <= h < Hi and 0 <= w < Wi
                    for hf in range(2):
                        piece = t + 4 * hf  # uint4 index inside the pixel row
                        v[(tap, r, hf)] = x_nhwc[st["pn"][r], h, w, 8 * piece:8 * piece + 8] if ok else np.zeros(8)
            st["v"] = v
            lanes.append(st)
        d = [[[0.0] * 4 for _ in range(32)] for _ in range(NT)]
        for ks in range(16):
            tap, s, hf = ks >> 2, ks & 3, (ks & 3) >> 1
            a_frag = []
            for lane in range(32):
                v = lanes[lane]["v"]
                lo, hi = (0, 2) if (s & 1) == 0 else (4, 6)  # .x/.y or .z/.w (pairs of channels)
                a_frag.append([v[(tap, 0, hf)][lo:lo + 2], v[(tap, 1, hf)][lo:lo + 2],
                               v[(tap, 0, hf)][hi:hi + 2], v[(tap, 1, hf)][hi:hi + 2]])
            for nt in range(NT):
                b_frag = [sb[(ks, nt, l)] for l in range(32)]
                mma_16816(a_frag, b_frag, d[nt])
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            st = lanes[lane]
            for nt in range(NT):
                for e in range(2):
                    co = nt * 8 + 2 * t + e
                    if co < CN:
                        for r in range(2):
                            if st["pv"][r]:
                                out[st["pn"][r], co, st["pi"][r], st["pj"][r]] = d[nt][lane][2 * r + e]
    return out


def narrow_in_emul(x, pre, pre_relu, wt, Ho, Wo, taps):
    """x planar [N][CN][Hi][Wi]; wt [tap][CN][64] -> out NHWC [N][Ho][Wo][64]."""
    N, CN, Hi, Wi = x.shape
    KS = (CN + 3) // 4
    n_taps = len(taps)
    out = np.zeros((N * Ho * Wo, 64))
    sb = {}
    for ks in range(KS):
        for nn in range(8):
            for l in range(32):
                g, t = l >> 2, l & 3
                co = 8 * (g >> 1) + 2 * nn + (g & 1) if nn < 4 else 32 + 8 * (g >> 1) + 2 * (nn - 4) + (g & 1)
                w = []
                for e in range(4):
                    k = 16 * ks + 2 * t + (e & 1) + (e >> 1) * 8
                    ci, tap = k >> 2, k & 3
                    w.append(wt[tap, ci, co] if (ci < CN and tap < n_taps) else 0.0)
                sb[(ks, nn, l)] = ((w[0], w[1]), (w[2], w[3]))
    npix = N * Ho * Wo
    tiles = (npix + 15) // 16
    for tile in range(tiles):
        d = [[[0.0] * 4 for _ in range(32)] for _ in range(8)]
        geo = []
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            rows = []
            for r in range(2):
                p = tile * 16 + g + 8 * r
                pv = p < npix
                pp = p if pv else 0
                q = pp // Wo
                rows.append((pv, pp, q // Ho, q - (q // Ho) * Ho, pp - q * Wo))
            geo.append(rows)
        for ks in range(KS):
            a_frag = []
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                regs = [None] * 4
                for cs in range(2):
                    ci = 4 * ks + 2 * cs + (t >> 1)
                    for r in range(2):
                        pv, pp, n, i, j = geo[lane][r]
                        pair = []
                        for e in range(2):
                            tap = 2 * (t & 1) + e
                            val = 0.0
                            if tap < n_taps:
                                dh, dw = taps[tap]
                                h, w = i + dh, j + dw
                                if pv and ci < CN and 0 <= h < Hi and 0 <= w < Wi:
                                    val = x[n, ci, h, w]
                                    if pre is not None:
                                        val = val * pre[ci] + pre[CN + ci]
                                        if pre_relu:
                                            val = max(val, 0.0)
                            pair.append(val)
                        regs[r + 2 * cs] = pair
                a_frag.append(regs)
            for nn in range(8):
                b_frag = [sb[(ks, nn, l)] for l in range(32)]
                mma_16816(a_frag, b_frag, d[nn])
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for r in range(2):
                pv, pp, n, i, j = geo[lane][r]
                if not pv:
                    continue
                o = [(d[nn][lane][2 * r], d[nn][lane][2 * r + 1]) for nn in range(8)]
                # dst[0] = piece t (channels 8t..8t+7), dst[4] = piece t+4
                out[pp, 8 * t:8 * t + 8] = [x_ for pr in o[0:4] for x_ in pr]
                out[pp, 8 * (t + 4):8 * (t + 4) + 8] = [x_ for pr in o[4:8] for x_ in pr]
    return out.reshape(N, Ho, Wo, 64)


def wtable(w, mode):
    """narrow_wtable_kernel: w OIHW [K][C][R][S]; mode 0 -> [tap][C][K], mode 1 -> [tap][K][C]."""
    K, C, R, S = w.shape
    t = w.permute(2, 3, 1, 0).reshape(R * S, C, K) if mode == 0 else w.permute(2, 3, 0, 1).reshape(R * S, K, C)
    return t.double().numpy()


def main():
    torch.manual_seed(0)
    for bch in (3, 6, 12):
        N, H, W = 2, 7, 9
        # enc7 forward: 64 -> bch, k2 p1
        x = torch.randn(N, 64, H, W, dtype=torch.float64)
        w7 = torch.randn(bch, 64, 2, 2, dtype=torch.float64)
        z = F.conv2d(x, w7, None, 1, 1)
        taps = [(r - 1, s - 1) for r in range(2) for s in range(2)]
        got = narrow_out_emul(x.permute(0, 2, 3, 1).numpy(), wtable(w7, 0), bch, H + 1, W + 1, taps)
        print("bch %d narrow_out fwd   max err %.2e" % (bch, np.abs(got - z.numpy()).max()))
        # enc7 dgrad (narrow_in flip): dz -> dx
        dz = torch.randn(z.shape, dtype=torch.float64)
        dx = torch.autograd.grad(F.conv2d(x.requires_grad_(True), w7, None, 1, 1), x, dz)[0]
        taps_f = [(1 - r, 1 - s) for r in range(2) for s in range(2)]
        # flip: narrow_wtable(w, C=bch(narrow) as K', K=64 as C', mode 1) -> [tap][narrow][wide]
        got = narrow_in_emul(dz.numpy(), None, False, wtable(w7, 1), H, W, taps_f)
        print("bch %d narrow_in  dgrad max err %.2e" % (bch, np.abs(got - dx.permute(0, 2, 3, 1).numpy()).max()))
        # dec2 forward: relu(bn(z)) -> 64, k2 p0
        zin = torch.randn(N, bch, H + 1, W + 1, dtype=torch.float64)
        sc, sh = torch.rand(bch, dtype=torch.float64) + 0.5, torch.randn(bch, dtype=torch.float64) * 0.3
        a = F.relu(zin * sc[None, :, None, None] + sh[None, :, None, None]).requires_grad_(True)
        w2 = torch.randn(64, bch, 2, 2, dtype=torch.float64)
        y = F.conv2d(a, w2)
        taps0 = [(r, s) for r in range(2) for s in range(2)]
        got = narrow_in_emul(zin.numpy(), torch.cat([sc, sh]).numpy(), True, wtable(w2, 0), H, W, taps0)
        print("bch %d narrow_in  fwd   max err %.2e" % (bch, np.abs(got - y.detach().permute(0, 2, 3, 1).numpy()).max()))
        # dec2 dgrad (narrow_out over dy): dy -> da
        dy = torch.randn(y.shape, dtype=torch.float64)
        da = torch.autograd.grad(y, a, dy)[0]
        taps_d = [(0 - r, 0 - s) for r in range(2) for s in range(2)]
        got = narrow_out_emul(dy.permute(0, 2, 3, 1).numpy(), wtable(w2, 1), bch, H + 1, W + 1, taps_d)
        print("bch %d narrow_out dgrad max err %.2e" % (bch, np.abs(got - da.numpy()).max()))


if __name__ == "__main__":
    main()
