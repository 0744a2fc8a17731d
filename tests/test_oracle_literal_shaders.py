"""The cleanup stages exist in the reference only as GLSL ("parity unpinned by any reference test").  The C++ oracle restates
them once; this file restates the shader TEXT a second time, independently and literally (one Python loop iteration per shader invocation,
the same expressions in the same order, float32 where the shader computes in float), and checks that both restatements agree on small
grids — including the quirks: the float index decode of voxel.glsl, the unclamped mask index of erodeGrid-comp.glsl:49, `globalCount` of the
clipped box, `isBoundary = bool(unmaskedBit(value, 15))` (every non-zero word passes), raw-word equality against tagged neighbours.

Shader sources followed (MeshFragments/Assets/Shaders/Compute/Fracturer/): voxel.glsl:6-19, voxelMask.glsl:4-17,
detectBoundaries-comp.glsl:18-43, erodeGrid-comp.glsl:26-59, removeIsolatedRegionsGrid-comp.glsl:16-39, undoMask-comp.glsl; host loops
RegularGrid.cpp:64-80 (detectBoundaries), :82-159 (erode: mask, activations, iteration order, final sweep), :1006-1015."""
import numpy as np
import pytest

f32 = np.float32
VOXEL_EMPTY, VOXEL_FREE, MASK_BOUNDARY_POSITION = 0, 1, 15


# ---- voxel.glsl / voxelMask.glsl
def get_position(index, dims):  # voxel.glsl:6-14 (float arithmetic as written; exact below 2^24 cells)
    X, Y, Z = dims
    x = f32(index) / f32(Y * Z)
    w = f32(index % (Y * Z))
    y = w / f32(Z)
    z = f32(int(w) % Z)
    return int(x), int(y), int(z)


def get_position_index(p, dims):  # voxel.glsl:16-19
    return p[0] * dims[1] * dims[2] + p[1] * dims[2] + p[2]


def masked_bit(v, pos):
    return (v | (1 << pos)) & 0xFFFF


def unmasked_bit(v, pos):
    return v & (~(1 << pos) & 0xFFFF)


def clamp3(p, dims):
    return tuple(min(max(p[k], 0), dims[k] - 1) for k in range(3))


# ---- detectBoundaries-comp.glsl:18-43, one invocation per cell, IN PLACE (the GPU's order is unspecified: any order must give the same result
# because neighbours are read with bit 15 cleared; the test runs two different orders)
def detect_boundaries_literal(grid, boundary_size, order):
    dims = grid.shape
    g = grid.reshape(-1).copy()
    for index in order:
        if g[index] <= VOXEL_FREE:
            continue
        gi = get_position(index, dims)
        lo = clamp3(tuple(c - boundary_size for c in gi), dims)
        hi = clamp3(tuple(c + boundary_size for c in gi), dims)
        boundary = False
        for x in range(lo[0], hi[0] + 1):
            for y in range(lo[1], hi[1] + 1):
                for z in range(lo[2], hi[2] + 1):
                    if boundary:
                        break
                    value = unmasked_bit(int(g[get_position_index((x, y, z), dims)]), MASK_BOUNDARY_POSITION)
                    boundary = boundary or (value > VOXEL_FREE and value != int(g[index]))
        if boundary:
            g[index] = masked_bit(int(g[index]), MASK_BOUNDARY_POSITION)
    return g.reshape(dims)


# ---- RegularGrid.cpp:84-122
def build_mask_literal(etype, size):
    if not size % 2:
        size += 1
    mask_size = size ** 3
    cc = int(np.floor(f32(size) / f32(2.0)))
    mask = np.zeros(mask_size, f32)
    activations = f32(0)
    if etype == 0:  # SQUARE
        mask[:] = 1
        activations = f32(mask_size)
    elif etype == 2:  # CROSS
        for a in range(size):
            mask[a * size * size + cc * size + cc] = 1
            mask[cc * size * size + a * size + cc] = 1
            mask[cc * size * size + cc * size + a] = 1
        activations = f32(1.0) / f32(3.0) * f32(mask_size)
    else:  # ELLIPSE
        for x in range(size):
            for y in range(size):
                for z in range(size):
                    d = np.sqrt(f32((x - cc) ** 2 + (y - cc) ** 2 + (z - cc) ** 2), dtype=f32)
                    if d < f32(cc) + np.finfo(f32).eps:
                        mask[x * size * size + y * size + z] = 1
                        activations = f32(activations + f32(1))
    return size, mask, f32(activations / f32(mask_size))


# ---- erodeGrid-comp.glsl:26-59, grid -> destGrid
def erode_pass_literal(grid, mask, size, noise, activations, prob, thr):
    dims = grid.shape
    g = grid.reshape(-1)
    dest = g.copy()  # :32
    k2 = int(np.floor(f32(size) / f32(2.0)))
    for index in range(g.size):
        own = int(g[index])
        is_boundary = bool(unmasked_bit(own, MASK_BOUNDARY_POSITION))  # :31, as written
        if own > VOXEL_FREE and is_boundary and noise[index % noise.size] < f32(prob):
            ind = get_position(index, dims)
            mx, mn = tuple(c + k2 for c in ind), tuple(c - k2 for c in ind)
            mxc, mnc = clamp3(mx, dims), clamp3(mn, dims)
            count = global_count = 0
            for x in range(mnc[0], mxc[0] + 1):
                for y in range(mnc[1], mxc[1] + 1):
                    for z in range(mnc[2], mxc[2] + 1):
                        same = f32(int(g[get_position_index((x, y, z), dims)]) == own)
                        count += int(same * mask[(x - mn[0]) * size * size + (y - mn[1]) * size + (z - mn[2])])
                        global_count += 1
            activation = f32(count) / f32(global_count)
            if activation < f32(activations) * f32(thr):
                dest[index] = VOXEL_EMPTY
    return dest.reshape(dims)


# ---- removeIsolatedRegionsGrid-comp.glsl:16-39 (in place on the GPU and therefore racy; DESIGN.md fixes "every invocation reads the grid as it
# was before the pass" — the only order-independent reading)
def sweep_literal(grid):
    dims = grid.shape
    g = grid.reshape(-1)
    out = g.copy()
    for index in range(g.size):
        count = -1
        ind = get_position(index, dims)
        mxc, mnc = clamp3(tuple(c + 1 for c in ind), dims), clamp3(tuple(c - 1 for c in ind), dims)
        for x in range(mnc[0], mxc[0] + 1):
            for y in range(mnc[1], mxc[1] + 1):
                for z in range(mnc[2], mxc[2] + 1):
                    count += int(g[get_position_index((x, y, z), dims)] == g[index])
        if count < 6:
            out[index] = VOXEL_EMPTY
    return out.reshape(dims)


def erode_literal(grid, noise, etype, size, iters, prob, thr):  # RegularGrid.cpp:126-156
    size, mask, activations = build_mask_literal(etype, size)
    g = grid.copy()
    for _ in range(iters):
        g = detect_boundaries_literal(g, 1, range(g.size))
        g = erode_pass_literal(g, mask, size, noise, activations, prob, thr)  # + copyGrid back (:149-152)
    return sweep_literal(g)


def _labelled(shape, seed, nlabels=4, fill=0.85):
    r = np.random.RandomState(seed)
    occ = r.rand(*shape) < fill
    lab = 2 + (r.randint(0, nlabels, (1, 1, shape[2])) + (np.arange(shape[0])[:, None, None] * nlabels // shape[0]) +
               (np.arange(shape[1])[None, :, None] * 2 // shape[1])) % nlabels
    g = np.where(occ, lab, 0).astype(np.uint16)
    g[r.rand(*shape) < 0.03] = 1  # a few FREE cells
    return g


@pytest.mark.parametrize("shape,seed", [((7, 6, 9), 1), ((4, 11, 5), 2), ((9, 3, 8), 3)])
def test_detect_boundaries_matches_the_shader_text_in_any_invocation_order(orc, shape, seed):
    g = _labelled(shape, seed)
    want = orc.detect_boundaries(g.copy(), 1)
    n = g.size
    assert np.array_equal(detect_boundaries_literal(g, 1, range(n)), want)
    assert np.array_equal(detect_boundaries_literal(g, 1, np.random.RandomState(seed).permutation(n)), want)
    # a second application on the tagged grid (RegularGrid::erode calls it once per iteration without undoMask in between)
    want2 = orc.detect_boundaries(want.copy(), 1)
    assert np.array_equal(detect_boundaries_literal(want, 1, range(n - 1, -1, -1)), want2)


@pytest.mark.parametrize("etype,size", [(0, 3), (1, 3), (2, 3), (1, 5), (0, 2)])
def test_erosion_mask_matches_the_host_code_text(orc, etype, size):
    k, mask, act = build_mask_literal(etype, size)
    m, a = orc.erode_mask(etype, size)
    assert m.shape == (k, k, k) and np.array_equal(m.reshape(-1), mask) and f32(a) == act


@pytest.mark.parametrize("shape,seed,etype,size,iters,prob,thr", [
    ((7, 6, 9), 11, 1, 3, 3, 0.5, 0.5),    # the reference's defaults: ELLIPSE 3, 3 iterations
    ((6, 8, 5), 12, 0, 3, 2, 0.7, 0.6),    # SQUARE
    ((5, 5, 10), 13, 2, 3, 2, 0.9, 0.95),  # CROSS with a threshold high enough to erode uniform neighbourhoods
    ((8, 7, 6), 14, 1, 5, 1, 0.8, 0.7),    # 5^3 mask: the unclamped mask index against the clipped loop bounds
])
def test_erode_matches_the_shader_text(orc, shape, seed, etype, size, iters, prob, thr):
    g = _labelled(shape, seed)
    noise = orc.Rng(seed).fill_noise(257)  # shorter than the grid: index % noiseBufferSize wraps
    want = orc.erode(g.copy(), noise, etype, size, iters, prob, thr, boundary_mode=0)
    got = erode_literal(g, noise, etype, size, iters, prob, thr)
    assert np.array_equal(got, want)
    assert (want != g).any()  # the case does erode something


@pytest.mark.parametrize("shape,seed", [((6, 7, 8), 21), ((3, 3, 3), 22), ((10, 2, 5), 23)])
def test_isolated_voxel_sweep_matches_the_shader_text(orc, shape, seed):
    g = _labelled(shape, seed, fill=0.7)
    g[::2, ::3, ::2] |= 0x8000  # tagged words are different words (raw compare)
    assert np.array_equal(sweep_literal(g), orc.remove_isolated_regions_grid(g.copy()))


def test_position_decode_of_the_shader_is_exact_on_these_sizes(orc):
    dims = (7, 6, 9)
    for index in range(7 * 6 * 9):
        assert get_position(index, dims) == np.unravel_index(index, dims)


@pytest.mark.parametrize("case", range(12))
def test_boundary_detection_after_the_first_iteration_is_a_no_op(orc, case):
    """RegularGrid::erode calls detectBoundaries(1) in every iteration (RegularGrid.cpp:133-135).  Tags are never cleared and erosion never
    creates a label, so from the second iteration on the pass cannot change the grid — which is why vf_erode launches it once
    (csrc/stencil.cu).  Checked on the literal transcription: the repeated pass is the identity, and the pipeline without it equals the
    oracle's pipeline with it, also when the grid arrives with tags already set."""
    rs = np.random.RandomState(100 + case)
    shape = tuple(int(v) for v in rs.randint(3, 9, 3))
    g = _labelled(shape, case, nlabels=int(rs.randint(2, 6)), fill=float(rs.uniform(0.5, 0.98)))
    if case % 3 == 0:
        g[rs.rand(*shape) < 0.2] |= 0x8000
    etype, size = int(rs.randint(0, 3)), int(rs.choice([3, 3, 5]))
    iters, prob, thr = int(rs.randint(2, 5)), float(rs.uniform(0.3, 1.0)), float(rs.uniform(0.3, 1.0))
    noise = orc.Rng(case).fill_noise(int(rs.randint(50, 400)))
    k, mask, act = build_mask_literal(etype, size)
    b = g.copy()
    for it in range(iters):
        tagged = detect_boundaries_literal(b, 1, range(b.size))
        if it == 0:
            b = tagged
        else:
            assert np.array_equal(tagged, b)
        b = erode_pass_literal(b, mask, k, noise, act, prob, thr)
    assert np.array_equal(sweep_literal(b), orc.erode(g.copy(), noise, etype, size, iters, prob, thr, boundary_mode=0))
