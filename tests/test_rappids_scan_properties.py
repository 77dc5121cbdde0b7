"""The order-independence arguments the planning pass of the batched RAPPIDS planner relies on
(agri-fly_b200/csrc/agf_rappids_plan.cuh: frame_pair / inflate, shrink_span, plan_vehicle), checked on the CPU with small
restatements of InflatePyramid's scans (Components/Components/DepthImagePlanner/DepthImagePlanner.cpp:519-599 spiral
expansion, :601-939 shrink updates of the four edge regions) -- pure Python, no device:

  * K iterations of the spiral expansion in which no side is blocked scan exactly the frame between the rectangle and
    the rectangle grown by K on every active side, and leave the same rectangle / flags / maxDepth as a single test of
    that frame; a frame with a blocker is refused and the line-by-line iterations decide.
  * the unblocked shrink updates of a span of an edge region equal one reduction over the pixels that trigger with the
    bounds at the start of the span.
  * rounds of up to four collision tests against copies of the pyramid list, committed in candidate order until the
    first result that moves the cost bound or grows the list, produce exactly the sequential plan (a toy deterministic
    test function stands in for IsCollisionFree).

The GPU tests (tests/test_rappids_gpu.py) compare the device code itself bit for bit with the line-by-line / per-pixel
variants and with the reference; these tests pin the ARGUMENT, so that it can be checked where there is no GPU.
"""
import numpy as np
import pytest

K_BUF = 2  # _pyramidSearchPixelBuffer (DepthImagePlanner.cpp:60)


# ------------------------------------------------------------------------------------------------
# spiral expansion
# ------------------------------------------------------------------------------------------------
def scan_line(pixels, min_pyr, ignore, max_depth):
    """One line of the expansion in scan order: (blocked, maxDepth) -- depths seen before the blocker are folded in."""
    for p in pixels:
        p = int(p)
        if p <= ignore:
            continue
        if p < min_pyr:
            return True, max_depth
        max_depth = min(max_depth, p)
    return False, max_depth


def one_iteration(img, st, min_pyr, ignore, edge_off):
    """Right, top, left, bottom as the reference does them (:521-597); st = [left, top, right, bottom, rf, tf, lf, bf, maxDepth]."""
    H, W = img.shape
    left, top, right, bottom, rf, tf, lf, bf, md = st
    if rf:
        if right < W - edge_off - 1:
            blk, md = scan_line(img[top:bottom + 1, right + 1], min_pyr, ignore, md)
            if blk:
                rf = False
            else:
                right += 1
        else:
            rf = False
    if tf:
        if top > edge_off:
            blk, md = scan_line(img[top - 1, left:right + 1], min_pyr, ignore, md)
            if blk:
                tf = False
            else:
                top -= 1
        else:
            tf = False
    if lf:
        if left > edge_off:
            blk, md = scan_line(img[top:bottom + 1, left - 1], min_pyr, ignore, md)
            if blk:
                lf = False
            else:
                left -= 1
        else:
            lf = False
    if bf:
        if bottom < H - edge_off - 1:
            blk, md = scan_line(img[bottom + 1, left:right + 1], min_pyr, ignore, md)
            if blk:
                bf = False
            else:
                bottom += 1
        else:
            bf = False
    return [left, top, right, bottom, rf, tf, lf, bf, md]


def expand_line_by_line(img, rect, min_pyr, ignore, edge_off):
    st = list(rect) + [True, True, True, True, 65535]
    iters = 0
    while st[4] or st[5] or st[6] or st[7]:
        st = one_iteration(img, st, min_pyr, ignore, edge_off)
        iters += 1
    return st, iters


def expand_with_frame_jumps(img, rect, min_pyr, ignore, edge_off, jump):
    """The device's driver loop (inflate): try a jump of K iterations; a refused jump is followed by up to K line-by-line
    iterations (fewer when a side stops)."""
    H, W = img.shape
    st = list(rect) + [True, True, True, True, 65535]
    cool = 0
    jumps = refused = 0
    while st[4] or st[5] or st[6] or st[7]:
        left, top, right, bottom, rf, tf, lf, bf, md = st
        if cool == 0:
            K = jump
            if rf:
                K = min(K, W - edge_off - 1 - right)
            if tf:
                K = min(K, top - edge_off)
            if lf:
                K = min(K, left - edge_off)
            if bf:
                K = min(K, H - edge_off - 1 - bottom)
            if K >= 2:
                r2, t2 = right + (K if rf else 0), top - (K if tf else 0)
                l2, b2 = left - (K if lf else 0), bottom + (K if bf else 0)
                # the frame as frame_pair takes it: rows above and below over the grown width, columns over the old height
                strips = []
                if tf:
                    strips.append(img[t2:top, l2:r2 + 1])
                if bf:
                    strips.append(img[bottom + 1:b2 + 1, l2:r2 + 1])
                if lf:
                    strips.append(img[top:bottom + 1, l2:left])
                if rf:
                    strips.append(img[top:bottom + 1, right + 1:r2 + 1])
                seen = np.concatenate([s.ravel() for s in strips]).astype(np.int64)
                seen = seen[seen > ignore]
                mn = int(seen.min()) if seen.size else 65535
                if mn >= min_pyr:  # no blocker anywhere in the frame
                    st = [l2, t2, r2, b2, rf, tf, lf, bf, min(md, mn)]
                    jumps += 1
                    continue
                refused += 1
                cool = K
        if cool > 0:
            cool -= 1
        flags0 = st[4:8]
        st = one_iteration(img, st, min_pyr, ignore, edge_off)
        if st[4:8] != flags0:
            cool = 0
    return st, jumps, refused


def random_scene(rng, H=96, W=128):
    img = np.full((H, W), 205, dtype=np.uint16)  # background
    band = rng.integers(0, H // 3)
    if band:  # a floor band whose depth falls towards the bottom
        img[H - band:, :] = (200 - np.arange(band) * rng.integers(1, 4))[:, None].clip(20, 205)
    for _ in range(rng.integers(0, 5)):
        y0, x0 = rng.integers(0, H - 4), rng.integers(0, W - 4)
        img[y0:y0 + rng.integers(2, H // 2), x0:x0 + rng.integers(2, W // 2)] = rng.integers(20, 160)
    if rng.random() < 0.5:  # pixels the planner does not see
        ys, xs = rng.integers(0, H, 40), rng.integers(0, W, 40)
        img[ys, xs] = rng.integers(0, 3, 40)
    return img


@pytest.mark.parametrize("jump", [2, 3, 8])
def test_frame_jumps_equal_the_line_by_line_expansion(jump):
    rng = np.random.default_rng(1234 + jump)
    total_jumps = total_refused = total_iters = 0
    for _ in range(400):
        img = random_scene(rng)
        H, W = img.shape
        edge_off = int(rng.integers(0, 6))
        r = int(rng.integers(1, 8))
        cx, cy = int(rng.integers(edge_off + r + 1, W - edge_off - r - 1)), int(rng.integers(edge_off + r + 1, H - edge_off - r - 1))
        rect = [cx - r, cy - r, cx + r, cy + r]
        min_pyr = int(rng.integers(10, 210))
        ignore = 2
        a, iters = expand_line_by_line(img, rect, min_pyr, ignore, edge_off)
        b, jumps, refused = expand_with_frame_jumps(img, rect, min_pyr, ignore, edge_off, jump)
        assert a == b, (rect, min_pyr, edge_off, a, b)
        total_jumps += jumps
        total_refused += refused
        total_iters += iters
    # the property is exercised: jumps are taken AND refused, and they replace most of the iterations
    assert total_jumps > 500 and total_refused > 100
    assert total_jumps * jump > 0.3 * total_iters


# ------------------------------------------------------------------------------------------------
# shrink updates of the four edge regions
# ------------------------------------------------------------------------------------------------
R_RIGHT, R_LEFT, R_TOP, R_BOTTOM = range(4)


def trigger(region, s, num, x, y, p):
    if region == R_RIGHT:
        return num > (x - s["r"]) * p
    if region == R_LEFT:
        return (s["l"] - x) * p < num
    if region == R_TOP:
        return (s["t"] - y) * p < num
    return num > (y - s["b"]) * p


def apply(region, s, num, x, y, p, x0, y0):
    """shrink_apply of an edge region (DepthImagePlanner.cpp:612-700 as restated on the device); False = cannot contain the point."""
    q = num // p
    rT, lT, tT, bT = x - q, x + q, y + q, y - q
    if region in (R_RIGHT, R_LEFT):
        blocked = (x0 > rT - K_BUF) if region == R_RIGHT else (x0 < lT + K_BUF)
        if not blocked:
            s["r" if region == R_RIGHT else "l"] = rT if region == R_RIGHT else lT
            return True
        no_top, no_bot = y0 < tT + K_BUF, y0 > bT - K_BUF
        if no_top and no_bot:
            return False
        if no_top:
            s["b"] = bT
        elif no_bot:
            s["t"] = tT
        else:
            u, d = tT - s["t"], s["b"] - bT
            if d > u:
                s["t"] = tT
            elif region == R_RIGHT:
                s["r"] = bT  # sic (DepthImagePlanner.cpp:648)
            else:
                s["b"] = bT
        return True
    blocked = (y0 < tT + K_BUF) if region == R_TOP else (y0 > bT - K_BUF)
    if not blocked:
        s["t" if region == R_TOP else "b"] = tT if region == R_TOP else bT
        return True
    no_right, no_left = x0 > rT - K_BUF, x0 < lT + K_BUF
    if no_right and no_left:
        return False
    if no_right:
        s["l"] = lT
    elif no_left:
        s["r"] = rT
    else:
        r, l = s["r"] - rT, lT - s["l"]
        if r > l:
            s["l"] = lT
        else:
            s["r"] = rT
    return True


def span_sequential(region, s, num, xs, ys, ps, x0, y0):
    s = dict(s)
    for x, y, p in zip(xs, ys, ps):
        if trigger(region, s, num, x, y, p):
            if not apply(region, s, num, x, y, p, x0, y0):
                return None
    return s


def span_folded(region, s, num, xs, ys, ps, x0, y0):
    """shrink_span's fast path; returns 'fallback' when a triggering pixel is blocked (the device then runs the sequential loop)."""
    s = dict(s)
    trig = [trigger(region, s, num, x, y, p) for x, y, p in zip(xs, ys, ps)]
    if not any(trig):
        return s
    cand, blocked = [], False
    for t, x, y, p in zip(trig, xs, ys, ps):
        if not t:
            continue
        q = num // p
        c = {R_RIGHT: x - q, R_LEFT: x + q, R_TOP: y + q, R_BOTTOM: y - q}[region]
        ref0 = x0 if region in (R_RIGHT, R_LEFT) else y0
        blocked |= (ref0 > c - K_BUF) if region in (R_RIGHT, R_BOTTOM) else (ref0 < c + K_BUF)
        cand.append(c)
    if blocked:
        return "fallback"
    key = {R_RIGHT: "r", R_LEFT: "l", R_TOP: "t", R_BOTTOM: "b"}[region]
    s[key] = min(cand) if region in (R_RIGHT, R_BOTTOM) else max(cand)
    return s


@pytest.mark.parametrize("region", [R_RIGHT, R_LEFT, R_TOP, R_BOTTOM])
def test_folded_shrink_span_equals_the_sequential_updates(region):
    rng = np.random.default_rng(77 + region)
    folded = fallback = 0
    for _ in range(4000):
        W, H = 320, 240
        num = int(rng.integers(100, 1500))
        x0, y0 = int(rng.integers(60, 260)), int(rng.integers(60, 180))
        s = dict(r=int(rng.integers(x0 + 3, W - 1)), l=int(rng.integers(0, x0 - 3)), t=int(rng.integers(0, y0 - 3)),
                 b=int(rng.integers(y0 + 3, H - 1)))
        n = int(rng.integers(1, 33))
        if region in (R_RIGHT, R_LEFT):  # column walk: one x, consecutive y
            x = int(rng.integers(x0 + 1, W)) if region == R_RIGHT else int(rng.integers(0, x0))
            ya = int(rng.integers(0, H - n))
            xs, ys = [x] * n, list(range(ya, ya + n))
        else:  # row walk: one y, consecutive x
            y = int(rng.integers(0, y0)) if region == R_TOP else int(rng.integers(y0 + 1, H))
            xa = int(rng.integers(0, W - n))
            xs, ys = list(range(xa, xa + n)), [y] * n
        base = int(rng.integers(3, 200))
        ps = [int(v) for v in np.clip(base + rng.integers(-40, 40, n), 3, 65535)]
        f = span_folded(region, s, num, xs, ys, ps, x0, y0)
        if f == "fallback":
            fallback += 1
            continue
        folded += 1
        assert f == span_sequential(region, s, num, xs, ys, ps, x0, y0), (region, s, num, xs, ys, ps, x0, y0)
    assert folded > 1000 and fallback > 50  # both paths occur


# ------------------------------------------------------------------------------------------------
# speculative planning of a long vehicle by a whole CTA (plan_vehicle, coop): rounds of up to four collision tests against
# copies of the pyramid list, committed in candidate order
# ------------------------------------------------------------------------------------------------
IN_FEASIBLE, CODE_VEL_OK = 0, 8


def toy_collision_test(cand, pyramids, rng_seed):
    """A deterministic function of (candidate, pyramid list) standing in for IsCollisionFree: may append pyramids to the
    list it is given (in place) and returns whether the candidate is free."""
    h = hash((rng_seed, cand, tuple(pyramids))) & 0xFFFFFFFF
    r = np.random.default_rng(h)
    for _ in range(int(r.integers(0, 3))):
        if r.random() < 0.25 and len(pyramids) < 32:
            pyramids.append(int(r.integers(0, 1 << 20)))
    return bool(r.random() < 0.08)


def plan_sequential(costs, codes, seed):
    best, pyr = float("inf"), []
    flags, cnt = [0] * len(costs), dict(cost=0, coll=0, vel=0, free=0)
    best_idx = -1
    for i, (c, code) in enumerate(zip(costs, codes)):
        if not c < best:
            continue
        f = 1
        cnt["cost"] += 1
        if (code & 7) == IN_FEASIBLE:
            f |= 2
            cnt["coll"] += 1
            if code & CODE_VEL_OK:
                f |= 4
                cnt["vel"] += 1
                if toy_collision_test(i, pyr, seed):
                    f |= 8
                    best, best_idx = c, i
                    cnt["free"] += 1
        flags[i] = f
    return flags, cnt, pyr, best_idx


def plan_speculative(costs, codes, seed, nw=4):
    """The device's loop: batches of 32 candidates; per round select up to nw tested candidates, test each against its own
    copy of the list, commit in order until the first result that moves the cost bound or grows the list."""
    best, truth = float("inf"), []
    flags, cnt = [0] * len(costs), dict(cost=0, coll=0, vel=0, free=0)
    best_idx, rounds, voided = -1, 0, 0
    for i0 in range(0, len(costs), 32):
        idx = list(range(i0, min(i0 + 32, len(costs))))
        pend = [i for i in idx if costs[i] < best]
        while pend:
            tested = lambda i: (codes[i] & 7) == IN_FEASIBLE and bool(codes[i] & CODE_VEL_OK)
            slots = [i for i in pend if costs[i] < best and tested(i)][:nw]
            results = []
            for i in slots:  # one warp each, its own copy of the list as it stood at the start of the round
                mine = list(truth)
                results.append((toy_collision_test(i, mine, seed), mine))
            rounds += 1
            s, stop, npyr0 = 0, False, len(truth)
            while pend and not stop:
                i = pend[0]
                if not costs[i] < best:
                    pend.pop(0)
                    continue
                if tested(i) and s >= len(slots):
                    break
                f = 1
                cnt["cost"] += 1
                if (codes[i] & 7) == IN_FEASIBLE:
                    f |= 2
                    cnt["coll"] += 1
                    if codes[i] & CODE_VEL_OK:
                        f |= 4
                        cnt["vel"] += 1
                        free, mine = results[s]
                        if free:
                            f |= 8
                            best, best_idx = costs[i], i
                            cnt["free"] += 1
                            stop = True
                        if len(mine) != npyr0:
                            truth = mine
                            stop = True
                        s += 1
                pend.pop(0)
                flags[i] = f
            voided += len(slots) - s
    return flags, cnt, truth, best_idx, rounds, voided


def test_speculative_rounds_commit_exactly_the_sequential_plan():
    rng = np.random.default_rng(5)
    voided = rounds = 0
    for trial in range(300):
        k = int(rng.integers(1, 200))
        costs = list(rng.normal(size=k) - np.linspace(0, rng.random() * 2, k))  # a drifting bound: many pass the cost check
        codes = [int(rng.choice([IN_FEASIBLE | CODE_VEL_OK] * 8 + [IN_FEASIBLE, 1, 2, 3])) for _ in range(k)]
        a = plan_sequential(costs, codes, trial)
        b = plan_speculative(costs, codes, trial)
        assert a == b[:4], trial
        rounds += b[4]
        voided += b[5]
    assert rounds > 1000 and voided > 100  # speculation both holds and is voided in this test
