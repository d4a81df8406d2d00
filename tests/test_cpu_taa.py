"""TAA resolve oracle (oracle/snk_taa.h, restating Core/res/shaders/TAA_Resolve.comp) against an independent numpy
restatement of the same shader and against known answers. CPU only."""
import numpy as np
import pytest

f32 = np.float32


def numpy_taa(cur, vel, hist, pvel):
    """Independent restatement (vectorised numpy, float32 ops in the shader's order)."""
    H, W = cur.shape[:2]
    cur32, hist32, vel32, pvel32 = (a.astype(f32) for a in (cur, hist, vel, pvel))
    ys, xs = np.mgrid[0:H, 0:W]
    hx = np.clip((xs.astype(f32) + f32(0.5)) / f32(W) + vel32[..., 0], f32(0), f32(1)).astype(f32)
    hy = np.clip((ys.astype(f32) + f32(0.5)) / f32(H) + vel32[..., 1], f32(0), f32(1)).astype(f32)
    mn = np.full((H, W, 3), 9999.0, f32); mx = np.full((H, W, 3), -9999.0, f32)
    m = np.zeros((H, W, 3), f32); m2 = np.zeros((H, W, 3), f32)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            ox, oy = np.clip(xs + dx, 0, W - 1), np.clip(ys + dy, 0, H - 1)
            col = np.maximum(cur32[oy, ox, :3], f32(0))
            mn = np.minimum(mn, col); mx = np.maximum(mx, col)
            m = (m + col).astype(f32); m2 = (m2 + (col * col).astype(f32)).astype(f32)
    inv = f32(1.0) / f32(9.0)
    mu = (m * inv).astype(f32)
    sigma = np.sqrt(np.abs(((m2 * inv).astype(f32) - (mu * mu).astype(f32)).astype(f32))).astype(f32)
    g = f32(0.55)
    minc, maxc = (mu - (g * sigma).astype(f32)).astype(f32), (mu + (g * sigma).astype(f32)).astype(f32)

    def sample(img, s, t):
        u = (s * f32(W) - f32(0.5)).astype(f32); v = (t * f32(H) - f32(0.5)).astype(f32)
        fu, fv = np.floor(u), np.floor(v)
        a, b = (u - fu).astype(f32)[..., None], (v - fv).astype(f32)[..., None]
        x0, y0 = fu.astype(np.int64) % W, fv.astype(np.int64) % H
        x1, y1 = (fu.astype(np.int64) + 1) % W, (fv.astype(np.int64) + 1) % H
        one = f32(1)
        w00, w10, w01, w11 = ((one - a) * (one - b)).astype(f32), (a * (one - b)).astype(f32), ((one - a) * b).astype(f32), (a * b).astype(f32)
        r = (img[y0, x0] * w00).astype(f32)
        r = (r + (img[y0, x1] * w10).astype(f32)).astype(f32)
        r = (r + (img[y1, x0] * w01).astype(f32)).astype(f32)
        return (r + (img[y1, x1] * w11).astype(f32)).astype(f32)

    h = np.maximum(sample(hist32, hx, hy)[..., :3], f32(0))
    p = np.concatenate([np.clip(mu, mn, mx), np.ones((H, W, 1), f32)], -1)
    q = np.concatenate([h, np.ones((H, W, 1), f32)], -1)
    r = (q - p).astype(f32)
    rmax, rmin = (maxc - p[..., :3]).astype(f32), (minc - p[..., :3]).astype(f32)
    eps = f32(0.00000001)
    with np.errstate(divide="ignore", invalid="ignore"):
        for c in range(3):
            s = (rmax[..., c] / r[..., c]).astype(f32)
            cond = r[..., c] > (rmax[..., c] + eps).astype(f32)
            r = np.where(cond[..., None], (r * s[..., None]).astype(f32), r)
        for c in range(3):
            s = (rmin[..., c] / r[..., c]).astype(f32)
            cond = r[..., c] < (rmin[..., c] - eps).astype(f32)
            r = np.where(cond[..., None], (r * s[..., None]).astype(f32), r)
    clipped = (p + r).astype(f32)[..., :3]
    pv = sample(pvel32, hx, hy)
    sx, sy = (pv[..., 0] + vel32[..., 0]).astype(f32), (pv[..., 1] + vel32[..., 1]).astype(f32)
    vl = np.sqrt(((sx * sx).astype(f32) + (sy * sy).astype(f32)).astype(f32)).astype(f32)
    dis = np.clip(((vl - f32(0.001)).astype(f32) * f32(10)).astype(f32), f32(0), f32(1))[..., None]

    def mix(a, b, t):
        return ((a * (f32(1) - t)).astype(f32) + (b * t).astype(f32)).astype(f32)
    fin = mix(mix(cur32[..., :3], clipped, f32(0.9)), mu, dis)
    out = cur.copy()
    gw, gh = (W // 8) * 8, (H // 8) * 8
    res = np.concatenate([fin, np.ones((H, W, 1), f32)], -1).astype(np.float16)
    out[:gh, :gw] = res[:gh, :gw]
    return out


def random_images(rng, H, W, motion=0.01):
    cur = (rng.random((H, W, 4)) * 4).astype(np.float16); cur[..., 3] = 1
    hist = (rng.random((H, W, 4)) * 4).astype(np.float16); hist[..., 3] = 1
    vel = ((rng.random((H, W, 2)) - 0.5) * 2 * motion).astype(np.float16)
    pvel = ((rng.random((H, W, 2)) - 0.5) * 2 * motion).astype(np.float16)
    return cur, vel, hist, pvel


@pytest.mark.parametrize("shape,motion", [((24, 32), 0.0), ((24, 32), 0.02), ((19, 45), 0.3), ((8, 8), 1.5)])
def test_oracle_matches_numpy_restatement(oracle, shape, motion):
    rng = np.random.default_rng(7 + shape[1])
    cur, vel, hist, pvel = random_images(rng, *shape, motion=motion)
    a = oracle.taa_resolve(cur, vel, hist, pvel)
    b = numpy_taa(cur, vel, hist, pvel)
    assert np.array_equal(a.view(np.uint16), b.view(np.uint16))


def test_known_answers(oracle):
    H, W = 16, 16
    # a constant frame with the same constant history is a fixed point (sigma = 0, clip box = the colour itself)
    c = np.zeros((H, W, 4), np.float16); c[...] = (0.5, 0.25, 2.0, 1.0)
    z = np.zeros((H, W, 2), np.float16)
    assert np.array_equal(oracle.taa_resolve(c, z, c, z), c)
    # a brighter history is clipped back to the neighbourhood box: the result stays the constant colour
    h = c.copy(); h[..., :3] = 60.0
    assert np.array_equal(oracle.taa_resolve(c, z, h, z), c)
    # fast motion (|v| >= 0.101 in uv) = full disocclusion: the output is the 3x3 mean of the clamped current frame
    rng = np.random.default_rng(3)
    cur, _, hist, _ = random_images(rng, H, W)
    v = np.zeros((H, W, 2), np.float16); v[..., 0] = 0.25
    out = oracle.taa_resolve(cur, v, hist, z).astype(np.float32)
    pad = np.pad(np.maximum(cur.astype(np.float32)[..., :3], 0), ((1, 1), (1, 1), (0, 0)), mode="edge")
    mean = sum(pad[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] for dx in (-1, 0, 1) for dy in (-1, 0, 1)) / 9.0
    assert np.allclose(out[..., :3], mean, rtol=2e-3, atol=2e-3)
    # sizes that are not a multiple of 8: the remainder keeps the rendered frame (dispatch truncation)
    cur, vel, hist, pvel = random_images(rng, 13, 21)
    out = oracle.taa_resolve(cur, vel, hist, pvel)
    assert np.array_equal(out[8:], cur[8:]) and np.array_equal(out[:, 16:], cur[:, 16:])
    assert not np.array_equal(out[:8, :16], cur[:8, :16])
