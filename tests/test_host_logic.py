"""CPU tests of the host-side mirror (data preparation, argument handling, sharding helpers)."""
import numpy as np
import pytest

import gpedm_b200 as G
from gpedm_b200 import batch, core, dist, fitgp, synth
from oracle import gpedm_oracle as O


def test_makelags_matches_reference_restatement(thetalog2pop):
    tl = thetalog2pop
    for E, tau in [(3, 1), (2, 2), (1, 3)]:
        got = G.makelags(data=tl, y="Abundance", pop="Population", E=E, tau=tau)
        ref = O.makelags_subrt(tl["Abundance"], tl["Population"], E, tau)
        assert list(got) == ["Abundance_%d" % (tau * (i + 1)) for i in range(E)]
        np.testing.assert_array_equal(np.column_stack(list(got.values())), ref)
    fore = G.makelags(data=tl, y="Abundance", pop="Population", time="Time", E=3, tau=1, forecast=True)
    assert list(fore) == ["Time", "Population", "Abundance_1", "Abundance_2", "Abundance_3"]
    np.testing.assert_array_equal(fore["Time"], [51, 51])
    ref = O.makelags_subrt(tl["Abundance"], tl["Population"], 3, 1, forecast=True)
    np.testing.assert_array_equal(np.column_stack([fore["Abundance_%d" % i] for i in (1, 2, 3)]), ref)


def test_makelags_multivariate_names_and_errors():
    y = np.arange(20.0).reshape(10, 2)
    got = G.makelags(y=y, E=2, tau=1, yname=["a", "b"])
    assert list(got) == ["a_1", "a_2", "b_1", "b_2"]
    assert np.isnan(got["a_2"][:2]).all() and got["a_2"][2] == 0.0
    with pytest.raises(ValueError):
        G.makelags(y=np.arange(3.0), E=3, tau=1)  # series length <= E*tau
    with pytest.raises(ValueError):
        G.makelags(y=np.arange(9.0), E=-1, tau=1)


def test_r2_helpers_match_oracle():
    rng = np.random.default_rng(1)
    obs, pred = rng.standard_normal(50), rng.standard_normal(50)
    obs[3] = np.nan
    pop = np.array(["a"] * 20 + ["b"] * 30, dtype=object)
    assert G.getR2(obs, pred) == pytest.approx(O.getR2(obs, pred), rel=1e-15)
    for t in ("centered", "scaled"):
        assert G.getR2pop(obs, pred, pop, t) == pytest.approx(O.getR2pop(obs, pred, pop, t), rel=1e-14)
    w = G.getR2pop(obs, pred, pop)
    wo = O.getR2pop(obs, pred, pop)
    assert w["R2pop"]["a"] == pytest.approx(wo["R2pop"]["a"]) and w["rmsepop"]["b"] == pytest.approx(wo["rmsepop"]["b"])
    with pytest.raises(ValueError):
        G.getR2pop(obs, pred, pop, "nope")


def test_pop_codes_and_rhomatrix_reordering():
    codes, levels = core.pop_codes(np.array(["z", "a", "z", "m"], dtype=object))
    assert levels == ["z", "a", "m"] and codes.tolist() == [0, 1, 0, 2] and codes.dtype == np.int32
    with pytest.raises(ValueError):
        core.pop_codes(["q"], levels)
    R = np.array([[1, .2, .3], [.2, 1, .4], [.3, .4, 1.]])
    out = core._rhomat_in_code_order(R, ["a", "m", "z"], levels)  # named a,m,z -> code order z,a,m
    np.testing.assert_array_equal(out, R[np.ix_([2, 0, 1], [2, 0, 1])])
    with pytest.raises(ValueError):
        core._rhomat_in_code_order(R, ["a", "m", "x"], levels)


def test_logit_and_default_start():
    assert core.logit(0.5) == 0.0
    assert core.logit(0.001, core.VEMIN, core.VEMAX) == pytest.approx(np.log((0.001 - 1e-4) / (4.9999 - 0.001)))
    t = batch.default_initparst(3, single_pop=False)
    np.testing.assert_allclose(t[:3], np.log(0.1))
    assert t[5] == 0.0  # logit(0.5, 0, 1), reference R/GPfunctions.R:424-426
    assert batch.default_initparst(2, True)[4] == 0.0


def test_fitgp_argument_errors_raise_before_touching_the_gpu():
    y = np.sin(np.arange(40.0))
    with pytest.raises(ValueError, match="Both E and tau"):
        G.fitGP(y=y, E=2)
    with pytest.raises(ValueError, match="x and/or E and tau"):
        G.fitGP(y=y)
    with pytest.raises(ValueError, match="either rhofixed or rhomatrix"):
        G.fitGP(y=y, E=1, tau=1, rhofixed=0.3, rhomatrix=np.eye(1))
    with pytest.raises(ValueError, match="fixedpars"):
        G.fitGP(y=y, E=2, tau=1, fixedpars=[0.1])
    with pytest.raises(ValueError, match="scaling"):
        G.fitGP(y=y, E=2, tau=1, scaling="weird")


def test_grid_problems_shapes():
    y = synth.ricker(300, 7)
    probs, meta = batch.grid_problems(y, [1, 2, 3], [1, 2])
    assert len(probs) == 6
    for p, m in zip(probs, meta):
        assert p["X"].shape == (300 - m["E"] * m["tau"], m["E"])
        assert not np.isnan(p["X"]).any() and not np.isnan(p["Y"]).any()
    Y, X = synth.embed(y, 3, 1)
    np.testing.assert_allclose(probs[4]["X"], X, rtol=1e-11, atol=1e-14)  # (E=3, tau=1)
    np.testing.assert_allclose(probs[4]["Y"], Y, rtol=1e-11, atol=1e-14)


def test_synthetic_generators_are_deterministic():
    np.testing.assert_array_equal(synth.ricker(50, 1004), synth.ricker(50, 1004))
    assert synth.hastings_powell(20, 5).shape == (20, 3)
    Y, X, pop = synth.hierarchical_thetalog(npop=3, n=40, E=2)
    assert X.shape == (3 * 38, 2) and len(np.unique(pop)) == 3


def test_shard_indices_partition():
    for n, w in [(40, 1), (40, 2), (40, 8), (5, 8)]:
        parts = [dist.shard_indices(n, w, r) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    parts = [dist.shard_indices(4, 2, r, costs=[10, 1, 1, 1]) for r in range(2)]
    assert parts[0] == [0] and parts[1] == [1, 2, 3]


def test_record_roundtrip():
    res = dict(status=0, iter=12, nevals=14, nllpost=-3.5, lp=-1.0, SS=2.0, logdet=0.5,
               pars=np.arange(5.0), parst=np.arange(5.0) - 1, grad=np.arange(5.0) * 2)
    i, back = dist.unpack_record(dist.pack_record(7, res))
    assert i == 7 and back["iter"] == 12 and back["nllpost"] == -3.5
    np.testing.assert_array_equal(back["grad"], res["grad"])


def _dag_tasks(nb, lauum_chunk=-1):
    import ctypes as C
    L = G.lib()
    n, nf = C.c_int(0), C.c_int(0)
    L.gpedm_debug_dag_tasks.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    assert L.gpedm_debug_dag_tasks(nb, lauum_chunk, None, 0, C.byref(n), C.byref(nf)) == 0
    buf = (C.c_int * (4 * n.value))()
    assert L.gpedm_debug_dag_tasks(nb, lauum_chunk, buf, n.value, C.byref(n), C.byref(nf)) == 0
    t = np.frombuffer(buf, dtype=np.int32).reshape(-1, 4)
    return [(int(a) & 15, int(i), int(j), int(w) & 1023, (int(w) >> 10) & 1023, (int(w) >> 20) & 1023, (int(w) >> 30) & 1)
            for a, i, j, w in t], nf.value


@pytest.mark.parametrize("nb,lch", [(1, -1), (2, -1), (3, -1), (4, -1), (5, 2), (8, -1), (8, 0), (13, 3), (32, -1), (32, 8), (64, -1)])
def test_dag_task_list_covers_every_tile_and_never_deadlocks(nb, lch):
    """The task list of the persistent factorisation kernel (csrc/host_handle.cuh build_dag_tasks), replayed on the
    CPU under the kernel's own rules (csrc/dag_factor.cuh): CTAs claim tasks in list order and wait inside a task
    for the ready flags of its operands.  Every tile's k range must be covered exactly once by consecutive pieces,
    and in-order claiming by a handful of CTAs (or by 148) must always make progress."""
    DIAG, POTRF, TRTRI, LAUUM = 0, 1, 2, 3
    tasks, nfactor = _dag_tasks(nb, lch)
    # ---- coverage
    cover = {}
    for ty, i, j, k0, k1, piece, last in tasks:
        c = cover.setdefault((ty, i, j), [0, 0, False])  # next k, next piece, closed
        assert not c[2] and k0 == c[0] and piece == c[1] and k1 >= k0
        c[0], c[1], c[2] = k1, piece + 1, bool(last)
    want = {}
    for c in range(nb):
        want[(DIAG, c, c)] = c
        for i in range(c + 1, nb):
            want[(POTRF, i, c)] = c
            want[(TRTRI, i, c)] = i - c
        for j in range(c + 1):
            want[(LAUUM, c, j)] = nb - c
    assert set(cover) == set(want)
    for key, nkt in want.items():
        assert cover[key][0] == nkt and cover[key][2], key
    assert sum(1 for t in tasks if t[0] != LAUUM) <= nfactor <= len(tasks)

    # ---- replay
    def ready(t, Lf, Xf, prog):
        ty, i, j, k0, k1, piece, last = t
        if piece > 0 and prog.get((ty, i, j), 0) < piece:
            return False
        if ty in (DIAG, POTRF):
            if not all(Lf[i][k] and Lf[j][k] for k in range(k0, k1)):
                return False
            return not (ty == POTRF and last) or Lf[j][j]
        if ty == TRTRI:
            if not all(Lf[i][j + k] and Xf[j + k][j] for k in range(k0, k1)):
                return False
            return not last or Lf[i][i]
        return all(Xf[i + k][i] and Xf[i + k][j] for k in range(k0, k1))

    for P in (3, 148):
        Lf = [[False] * nb for _ in range(nb)]
        Xf = [[False] * nb for _ in range(nb)]
        prog, nxt, held, done = {}, 0, [], 0
        while done < len(tasks):
            while len(held) < P and nxt < len(tasks):
                held.append(tasks[nxt])
                nxt += 1
            run = [t for t in held if ready(t, Lf, Xf, prog)]
            assert run, "deadlock with %d CTAs at task %d of %d (nb=%d)" % (P, nxt, len(tasks), nb)
            for t in run:
                ty, i, j, k0, k1, piece, last = t
                held.remove(t)
                done += 1
                if not last:
                    prog[(ty, i, j)] = piece + 1
                elif ty == DIAG:
                    Lf[i][i] = Xf[i][i] = True
                elif ty == POTRF:
                    Lf[i][j] = True
                elif ty == TRTRI:
                    Xf[i][j] = True
