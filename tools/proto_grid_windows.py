"""Prototype (NumPy) of the closed-form anchor windows used by the grid encode driver.

For every GT box and every (feature map, box shape) grid of the anchor table it computes the rectangle of (row, col)
cells that can hold an anchor with IoU >= t and checks, against the full IoU matrix of the oracle, that no pair with
IoU >= t lies outside its rectangle; prints how many cells per image the rectangles hold.
Run: python tools/proto_grid_windows.py [std|stress] [images]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import ssd_oracle as orc  # noqa: E402

F32 = np.float32


def detect_grids(anc):
    """Split the (A,4) table into maps (off, H, W, nb) with x extents independent of the row and y extents of the column."""
    A = anc.shape[0]
    grids = []
    off = 0
    while off < A:
        rem = A - off
        found = None
        # nb: smallest n with the same y extents at off+n (next column, same shape)
        for nb in range(1, min(rem, 64)):
            if anc[off + nb, 1] == anc[off, 1] and anc[off + nb, 3] == anc[off, 3]:
                # W: smallest c with the same x extents at off + c*nb (next row)
                W = None
                for c in range(2, rem // nb + 1):
                    i = off + c * nb
                    if i >= A:
                        break
                    if anc[i, 0] == anc[off, 0] and anc[i, 2] == anc[off, 2]:
                        W = c
                        break
                    if not (anc[i, 1] == anc[off, 1] and anc[i, 3] == anc[off, 3]):
                        break
                if W is None:
                    continue
                L = W * nb
                H = 1
                while off + (H + 1) * L <= A:
                    r = anc[off + H * L: off + (H + 1) * L]
                    if np.array_equal(r[:, [0, 2]], anc[off: off + L][:, [0, 2]]) and np.all(
                            r.reshape(W, nb, 4)[:, :, [1, 3]] == r.reshape(W, nb, 4)[0:1, :, [1, 3]]):
                        H += 1
                    else:
                        break
                found = (off, H, W, nb)
                break
        if found is None:
            return None
        grids.append(found)
        off += found[1] * found[2] * found[3]
    return grids


def windows_for_image(anc, grids, gt, t_g):
    """Returns list of (g, off, H, W, nb, k, r0, r1, c0, c1) rectangles (inclusive)."""
    out = []
    for (off, H, W, nb) in grids:
        tab = anc[off: off + H * W * nb].reshape(H, W, nb, 4)
        for k in range(nb):
            X0, X1 = tab[0, :, k, 0].astype(np.float64), tab[0, :, k, 2].astype(np.float64)
            Y0, Y1 = tab[:, 0, k, 1].astype(np.float64), tab[:, 0, k, 3].astype(np.float64)
            # linear model of the centre + constant half extent
            cx = 0.5 * (X0 + X1)
            cy = 0.5 * (Y0 + Y1)
            hx = 0.5 * (X1 - X0).max()
            hy = 0.5 * (Y1 - Y0).max()
            dx = (cx[-1] - cx[0]) / (W - 1) if W > 1 else 1.0
            dy = (cy[-1] - cy[0]) / (H - 1) if H > 1 else 1.0
            ex = np.abs(cx - (cx[0] + dx * np.arange(W))).max() + np.abs(0.5 * (X1 - X0) - hx).max()
            ey = np.abs(cy - (cy[0] + dy * np.arange(H))).max() + np.abs(0.5 * (Y1 - Y0) - hy).max()
            aw, ah = 2 * hx + 1, 2 * hy + 1
            amin = ((Y1 - Y0 + 1).min()) * ((X1 - X0 + 1).min())
            for g in range(gt.shape[0]):
                gx0, gy0, gx1, gy1 = [float(v) for v in gt[g]]
                gw, gh = gx1 - gx0 + 1, gy1 - gy0 + 1
                ga = gw * gh
                t = float(t_g[g])
                imin = t * (amin + ga) / (1 + t) * (1 - 1e-5)
                iwmax, ihmax = min(aw, gw), min(ah, gh)
                if imin > iwmax * ihmax * (1 + 1e-5):
                    continue
                nw = imin / ihmax - 1e-3
                nh = imin / iwmax - 1e-3
                px, py = ex + 1e-3, ey + 1e-3
                lo = gx0 + nw - 1 - hx - px
                hi = gx1 - nw + 1 + hx + px
                c0 = int(np.ceil((lo - cx[0]) / dx - 1e-6)) if W > 1 else 0
                c1 = int(np.floor((hi - cx[0]) / dx + 1e-6)) if W > 1 else (0 if lo <= cx[0] <= hi else -1)
                lo = gy0 + nh - 1 - hy - py
                hi = gy1 - nh + 1 + hy + py
                r0 = int(np.ceil((lo - cy[0]) / dy - 1e-6)) if H > 1 else 0
                r1 = int(np.floor((hi - cy[0]) / dy + 1e-6)) if H > 1 else (0 if lo <= cy[0] <= hi else -1)
                c0, c1, r0, r1 = max(c0, 0), min(c1, W - 1), max(r0, 0), min(r1, H - 1)
                if c0 <= c1 and r0 <= r1:
                    out.append((g, off, H, W, nb, k, r0, r1, c0, c1))
    return out


def seeds_for_image(anc, grids, gt):
    """Exact IoU of the nearest-centre anchor of every (map, shape) -> lower bound of max_a iou[:, g]."""
    G = gt.shape[0]
    best = np.zeros(G, F32)
    for (off, H, W, nb) in grids:
        tab = anc[off: off + H * W * nb].reshape(H, W, nb, 4)
        for k in range(nb):
            cx = 0.5 * (tab[0, :, k, 0] + tab[0, :, k, 2]).astype(np.float64)
            cy = 0.5 * (tab[:, 0, k, 1] + tab[:, 0, k, 3]).astype(np.float64)
            dx = (cx[-1] - cx[0]) / (W - 1) if W > 1 else 1.0
            dy = (cy[-1] - cy[0]) / (H - 1) if H > 1 else 1.0
            for g in range(G):
                gcx, gcy = 0.5 * (gt[g, 0] + gt[g, 2]), 0.5 * (gt[g, 1] + gt[g, 3])
                c = int(min(max(round((gcx - cx[0]) / dx), 0), W - 1))
                r = int(min(max(round((gcy - cy[0]) / dy), 0), H - 1))
                a = off + (r * W + c) * nb + k
                iou = orc.iou_matrix(anc[a: a + 1], gt[g: g + 1])[0, 0]
                best[g] = max(best[g], iou)
    return best


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "std"
    n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rng = np.random.default_rng(1993)
    if cfg == "std":
        anc = orc.anchors_generate(**orc.STD_ANCHOR_CFG)
        labels, boxes, counts = orc.synth_gt(rng, n_img, 16)
    else:
        anc = orc.anchors_generate(**orc.STRESS_ANCHOR_CFG)
        labels, boxes, counts = orc.synth_gt(rng, n_img, 256, image_hw=(1000, 1280))
    grids = detect_grids(anc)
    print("grids", grids)
    thr = 0.5
    tot_cells, tot_pos, tot_win = 0, 0, 0
    for b in range(n_img):
        gt = boxes[b]
        iou = orc.iou_matrix(anc, gt)
        seed = seeds_for_image(anc, grids, gt)
        true_best = iou.max(axis=0)
        assert np.all(seed <= true_best)
        t_g = np.minimum(F32(thr), seed)
        wins = windows_for_image(anc, grids, gt, t_g)
        covered = np.zeros(iou.shape, bool)
        cells = 0
        for (g, off, H, W, nb, k, r0, r1, c0, c1) in wins:
            rr, cc = np.meshgrid(np.arange(r0, r1 + 1), np.arange(c0, c1 + 1), indexing="ij")
            a = off + (rr * W + cc) * nb + k
            covered[a.ravel(), g] = True
            cells += a.size
        need = iou >= t_g[None, :]
        need &= iou > 0
        miss = need & ~covered
        assert not miss.any(), (b, np.argwhere(miss)[:5])
        tot_cells += cells
        tot_win += len(wins)
        tot_pos += int((iou > thr).sum())
        if b < 3:
            print(f"image {b}: windows {len(wins)} cells {cells} pairs>thr {(iou > thr).sum()} "
                  f"seed==best {(seed == true_best).sum()}/{len(seed)} seed<thr {(seed < thr).sum()}")
    print(f"avg windows {tot_win / n_img:.1f} cells {tot_cells / n_img:.1f} pairs>thr {tot_pos / n_img:.1f}")


if __name__ == "__main__":
    main()
