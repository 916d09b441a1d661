"""A CPU model of the aggregation kernel's intra-CTA schedule (s2p_b200/csrc/agg_kernel.cuh, run_band): which pixel
every warp writes to its shared-memory ring and reads from its predecessor's ring at every lock-step, and where the CTA
barriers fall.  Between two barriers the warps are unordered, so a value must be written in an earlier barrier
interval than the one it is read in, and a ring slot may only be overwritten in a later interval than its last read.
The model checks the shipped configurations and shows that it rejects the ones the kernel must not use."""
import itertools

import pytest

K_NWC = 8          # warps per CTA, two scanlines each


def simulate(skew, lead, sync2, ring, n_i=45):
    """-> list of hazards.  skew: 1, or 2 when the neighbour E is used (then lead = 1: the newest pixel of the previous
    scanline needed at position i is i + 1).  sync2: barrier after every second step, warps skewed by one more pixel."""
    wsk = 2 * skew + 1 if sync2 else 2 * skew
    nsteps = n_i + (K_NWC - 1) * wsk + skew + 2

    def interval(t):           # index of the barrier interval step t belongs to
        return (t + 1) // 2 if sync2 else t        # barriers after odd steps only | after every step

    written, reads = {}, {}    # (warp, pixel) -> interval of the write ; (warp, pixel) -> interval of the last read by warp+1
    for t in range(nsteps):
        for k in range(K_NWC):
            i_a = t - k * wsk
            i_b = i_a - skew
            if 0 <= i_b < n_i:                                   # scanline B's result goes to this warp's ring
                written[(k, i_b)] = interval(t)
            if k >= 1 and 0 <= i_a < n_i:                        # scanline A reads the previous warp's ring
                need = []
                if lead:
                    if i_a == 0:
                        need.append(0)
                    if i_a + 1 < n_i:
                        need.append(i_a + 1)
                else:
                    need.append(i_a)
                for j in need:
                    reads[(k - 1, j)] = max(reads.get((k - 1, j), -1), interval(t))
    hazards = []
    for (k, j), r in reads.items():
        w = written.get((k, j))
        if w is None or not w < r:
            hazards.append(("read-before-write", k, j, w, r))
        over = written.get((k, j + ring))                        # the write that reuses pixel j's slot
        if over is not None and not r < over:
            hazards.append(("overwritten-before-read", k, j, r, over))
    return hazards


@pytest.mark.parametrize("skew,lead", [(1, 0), (2, 1)])
def test_shipped_schedules_are_hazard_free(skew, lead):
    assert simulate(skew, lead, sync2=False, ring=4) == []       # every slab width: one barrier per step, ring of 4
    assert simulate(skew, lead, sync2=True, ring=8) == []        # 5..8 labels per lane: one barrier per two steps, ring of 8


@pytest.mark.parametrize("skew,lead", [(1, 0), (2, 1)])
def test_model_rejects_unsafe_variants(skew, lead):
    # barrier every second step WITHOUT the extra pixel of skew between warps: the reader can overtake the writer
    def no_extra_skew(skew, lead):
        wsk = 2 * skew
        bad = []
        for t, k in itertools.product(range(60), range(1, K_NWC)):
            i_a = t - k * wsk
            j = i_a + lead
            t_w = j + (k - 1) * wsk + skew                       # step at which warp k-1 wrote pixel j
            if 0 <= i_a < 40 and (t_w + 1) // 2 >= (t + 1) // 2:
                bad.append((t, k))
        return bad
    assert no_extra_skew(skew, lead)
    if lead:     # with the neighbour E, pixel 0 is read three steps after it was written: a ring of 4 is then too short
        assert any(h[0] == "overwritten-before-read" for h in simulate(skew, lead, sync2=True, ring=4))
    assert any(h[0] == "overwritten-before-read" for h in simulate(skew, lead, sync2=False, ring=1))


def simulate_chunked(skew, use_e, ring=4, n_i=40, warps=16):
    """The experimental chunk-skipping kernel (agg_chunked.cuh): one scanline per warp, a barrier after every step,
    warp k at pixel t - k*skew; it reads pixels i-1, i (and i+1 with the neighbour E) of the previous warp's ring and
    pixel i-1 of its own."""
    written, reads = {}, {}
    for t in range(n_i + (warps - 1) * skew):
        for k in range(warps):
            i = t - k * skew
            if not 0 <= i < n_i:
                continue
            written[(k, i)] = t
            if 0 < i < n_i - 1:                      # border pixels read nothing
                reads[(k, i - 1)] = max(reads.get((k, i - 1), -1), t)            # own previous pixel
                if k >= 1:
                    for j in (i - 1, i) + ((i + 1,) if use_e else ()):
                        reads[(k - 1, j)] = max(reads.get((k - 1, j), -1), t)
    hazards = []
    for (k, j), r in reads.items():
        w = written.get((k, j))
        if w is None or not w < r:
            hazards.append(("read-before-write", k, j, w, r))
        over = written.get((k, j + ring))
        if over is not None and not r < over:
            hazards.append(("overwritten-before-read", k, j, r, over))
    return hazards


@pytest.mark.parametrize("skew,use_e", [(1, False), (2, True)])
def test_chunked_kernel_schedule(skew, use_e):
    assert simulate_chunked(skew, use_e, ring=4) == []
    assert simulate_chunked(skew, use_e, ring=2) != []          # the model has teeth
    if use_e:
        assert simulate_chunked(1, True) != []                  # E needs the two-pixel skew
