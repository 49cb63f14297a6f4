"""CPU restatements of index arithmetic that the CUDA kernels rely on (no GPU, no library call): a change of the
device code that breaks one of these identities would only show on a B200, so the arithmetic is pinned here too."""
import numpy as np
import pytest


def _hterms_staged_reduction(D, n_glag, s, J, Kp):
    """The folded-lag reduction of tc_hterms_kernel (cmfpy_b200/csrc/tc_strict_kernels.cuh, `p.staged`), lane by lane.

    D[row][col]: the 128 x 256 accumulator of one work item, row = g * 32 + d * Kp + k (lag group, lag inside a
    virtual lag, component).  Returns R[k][u], u = 0 .. 256 + hd - 1."""
    hd = (n_glag - 1) * s * J + s - 1
    U = 256 + hd
    Up = U | 1
    R = np.zeros((Kp, Up))
    sJ = s * J
    for c in range(4):                                   # 32 accumulator columns of each half at a time
        St = [D[:, h * 128 + c * 32: h * 128 + c * 32 + 32] for h in (0, 1)]        # the staging tiles
        for k in range(Kp):                              # warp e takes k = e, e + 8, ...
            for r in range(32):                          # lane
                shift_g = (n_glag - 1) * sJ + (s - 1)
                for g in range(n_glag):
                    shift = shift_g
                    for d in range(s):
                        row = g * 32 + d * Kp + k
                        j0 = (r - shift) & 31
                        u = c * 32 + r + ((shift - r + 31) & ~31)
                        R[k, u] += St[0][row, j0]
                        R[k, u + 128] += St[1][row, j0]
                        shift -= 1
                    shift_g -= sJ
    return R[:, :U], hd


@pytest.mark.parametrize("Kp,J", [(8, 2), (8, 1), (8, 16), (16, 3), (16, 1)])
def test_hterms_staged_reduction_places_every_accumulator_once(Kp, J):
    """Every accumulator element D[(g, d, k)][col] must land exactly once on R[k][col + shift(g, d)],
    shift = (n_glag - 1 - g) s J + (s - 1 - d): the definition of the lag-group reduction (the four lag groups of the
    H-terms kernel share the 128 MMA rows; reference tensor_transconv, cmfpy/common.py:61-86)."""
    s, n_glag = 32 // Kp, 4
    rng = np.random.default_rng(Kp * 100 + J)
    D = rng.random((128, 256))
    R, hd = _hterms_staged_reduction(D, n_glag, s, J, Kp)
    want = np.zeros((Kp, 256 + hd))
    for g in range(n_glag):
        for d in range(s):
            shift = (n_glag - 1 - g) * s * J + (s - 1 - d)
            for k in range(Kp):
                want[k, shift:shift + 256] += D[g * 32 + d * Kp + k]
    assert np.allclose(R, want, rtol=0, atol=1e-12)


@pytest.mark.parametrize("limit,tau0,want", [(100, 0, 32), (100, 96, 4), (100, 100, 0), (100, 128, 0), (0, 0, 0),
                                             (2 ** 40, 2 ** 40 - 5, 5)])
def test_rows_below(limit, tau0, want):
    """rows_below (tc_kernels.cuh): how many of the 32 rows tau0 .. tau0 + 31 lie below `limit` - the per-block
    count that replaced the per-element 64-bit `tau < t_own` / `tau < t_valid` tests of the K1 epilogues."""
    d = limit - tau0
    got = 0 if d <= 0 else (32 if d >= 32 else d)
    assert got == want == sum(1 for j in range(32) if tau0 + j < limit)


@pytest.mark.parametrize("no", [1, 2, 17, 31, 32])
def test_clamped_row_walk_reads_only_valid_rows(no):
    """recon_load_x (tc_kernels.cuh): the running pointer advances only while the next row is below `no`, so rows
    >= no re-read row no - 1 (valid memory) and are zeroed afterwards."""
    rows, row = [], 0
    for j in range(32):
        rows.append(row)
        row += 1 if j + 1 < no else 0
    assert max(rows) == no - 1 and rows[:no] == list(range(no)) and all(r == no - 1 for r in rows[no:])


@pytest.mark.parametrize("n_src,n_split,n_chunks_n,n_tiles", [(1, 1, 8, 5), (2, 1, 32, 3), (1, 2, 32, 7), (2, 2, 16, 4),
                                                              (2, 4, 16, 3)])
def test_hterms_item_split_covers_every_chunk_once(n_src, n_split, n_chunks_n, n_tiles):
    """Work items of tc_hterms_kernel with HTermsParams::n_split (short shards): item -> (time tile, source, split);
    the producer's feature chunk is (item % n_split) * cpi + nc, the epilogue's output slot source * n_split + split.
    Every (tile, source, chunk) must be contracted exactly once, and the slots of one source must be the n_split
    consecutive partial outputs that the host sums (tc_path.cuh h_terms)."""
    cpi = n_chunks_n // n_split
    per_tile = n_src * n_split
    seen = {}
    for item in range(n_tiles * per_tile):
        tile, rem = divmod(item, per_tile)
        src_w, nc0_w = rem // n_split, (rem % n_split) * cpi          # window loads
        nc0_a = (item % n_split) * cpi                                # W stage loads
        slot = item % per_tile                                        # epilogue
        assert nc0_w == nc0_a and slot // n_split == src_w
        for nc in range(cpi):
            key = (tile, src_w, nc0_w + nc)
            assert key not in seen
            seen[key] = slot
    assert len(seen) == n_tiles * n_src * n_chunks_n
    for (tile, src, chunk), slot in seen.items():
        assert src * n_split <= slot < (src + 1) * n_split and slot - src * n_split == chunk // cpi
