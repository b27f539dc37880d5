"""Design check for DESIGN.md 9 (1): the steady-state column of the tile kernel as an IN-PLACE butterfly with fixed read positions.

Today (tile_fast.h: column_fast) the projection index lists the local reads by age — bit 0 is the read that ends, the read that starts
becomes the top bit — so every column moves every entry (out index = in index >> 1 | new bit << top): a round trip through shared
memory per column.  If a read keeps ONE bit position for its whole life and the read that starts inherits the position of the read that
ends, a column only combines the two entries that differ in that position:

    out[y | t << q] = min_b ( cost_k(y, b, t) + in[y | b << q] ),   b = side of the read that ends, t = side of the read that starts,

an in-place butterfly on bit q with q changing from column to column.  This test runs both layouts over many columns in numpy (values
below 2^28 like the tile path; small weights so that ties are frequent) and checks that they describe the same DP: equal values for
equal read-side assignments and the same winning candidate under the reference's tie-break (candidate 1 wins iff
v1 < v0 + par, par = parity of all other reads' sides + the global reads' constant: tile_fast.h, pedigreedptable.cpp:293-296)."""
import numpy as np


def column_cost(u, k12):
    return np.minimum(u, (k12 - u) & 0xFFFFFFFF)  # u32: min(K2 + E, K1 - E)


def canonical_column(vals, w_stay, w_end, w_new, k2, k12, cg):
    """vals[x]: x lists the L local reads by age (bit 0 = the read that ends).  Returns (out[o], pick[o]); o = staying reads by age, new read on top."""
    L = int(np.log2(vals.size))
    o = np.arange(vals.size, dtype=np.int64)
    low = o & ((1 << (L - 1)) - 1)          # sides of the L - 1 reads that stay (their age order is kept)
    t = o >> (L - 1)                        # side of the read that starts
    e_stay = sum(((low >> j) & 1) * w_stay[j] for j in range(L - 1))
    par = ((np.bitwise_count(o.astype(np.uint64)).astype(np.int64) + cg) & 1)
    v = []
    for b in (0, 1):
        u = (k2 + e_stay + t * w_new + b * w_end) & 0xFFFFFFFF
        v.append(column_cost(u, k12) + vals[(low << 1) | b])
    pick = v[1] < v[0] + par
    return np.minimum(v[0], v[1]), pick


def butterfly_column(vals, pos_weight, q, w_end, w_new, k2, k12, cg):
    """vals[x]: bit p of x = side of the read at POSITION p; the read that ends sits at position q and the read that starts takes it over.
    pos_weight[p]: weight of the read at position p (entry q is ignored).  In place: returns (vals', pick') over the same index space."""
    L = int(np.log2(vals.size))
    x = np.arange(vals.size, dtype=np.int64)
    y = x & ~(1 << q)
    t = (x >> q) & 1                        # the output's bit q is the new read's side
    e_other = sum(((y >> p) & 1) * pos_weight[p] for p in range(L) if p != q)
    par = ((np.bitwise_count(y.astype(np.uint64)).astype(np.int64) + t + cg) & 1)
    v = []
    for b in (0, 1):
        u = (k2 + e_other + t * w_new + b * w_end) & 0xFFFFFFFF
        v.append(column_cost(u, k12) + vals[y | (b << q)])
    pick = v[1] < v[0] + par
    return np.minimum(v[0], v[1]), pick


def test_in_place_butterfly_equals_the_shifting_layout():
    rng = np.random.default_rng(20250924)
    L = 10
    for trial in range(4):
        # reads 0 .. L-1 are active, read r at age r (canonical) and at position perm[r] (fixed layout)
        ages = list(range(L))                                  # ages[j] = read at canonical bit j
        position = {r: int(p) for r, p in enumerate(rng.permutation(L))}
        weight = {r: int(rng.integers(-3, 4)) for r in range(L)}
        canon = rng.integers(0, 50, 1 << L).astype(np.int64)   # canon[x]: sides by age
        # the same function in the fixed layout
        idx = np.arange(1 << L, dtype=np.int64)
        def to_fixed(reads_by_age, pos):
            out = np.zeros_like(idx)
            for j, r in enumerate(reads_by_age):
                out |= ((idx >> j) & 1) << pos[r]
            return out                                         # fixed index of the canonical index
        fixed = np.zeros_like(canon)
        fixed[to_fixed(ages, position)] = canon
        next_read = L
        for col in range(60):
            ending, starting = ages[0], next_read
            next_read += 1
            weight[starting] = int(rng.integers(-3, 4))
            # the tile path's cost form: cost = min(K2 + E, K1 - E), E = sum of the signed weights of the reads on side 1;
            # K2 = sum of |w| over negative weights (+ the global reads' part), K1 the same over positive ones: both terms stay >= 0
            active = ages + [starting]
            g2, g1 = int(rng.integers(0, 10)), int(rng.integers(0, 10))
            k2 = g2 + sum(-weight[r] for r in active if weight[r] < 0)
            k12 = k2 + g1 + sum(weight[r] for r in active if weight[r] > 0)
            cg = int(rng.integers(0, 2))
            out_c, pick_c = canonical_column(canon, [weight[r] for r in ages[1:]], weight[ending], weight[starting], k2, k12, cg)
            q = position[ending]
            pos_weight = [0] * L
            for r in ages[1:]:
                pos_weight[position[r]] = weight[r]
            out_f, pick_f = butterfly_column(fixed, pos_weight, q, weight[ending], weight[starting], k2, k12, cg)
            ages = ages[1:] + [starting]
            position[starting] = q
            del position[ending]
            m = to_fixed(ages, position)
            assert np.array_equal(out_f[m], out_c), (trial, col)
            assert np.array_equal(pick_f[m], pick_c), (trial, col)
            canon, fixed = out_c, out_f
            assert int(canon.max()) < (1 << 28)
