#!/usr/bin/env python3
"""Lock-step replay of a golden fixture (or a seeded oracle run) by the oracle and the CUDA
model, comparing after every sub-move; on the first Gibbs mismatch prints the first visit
whose decision differs together with the oracle's weights for it.  Debug aid (GPU box)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))

from helpers import Golden, golden_names  # noqa: E402
from oracle.crp_oracle import (OracleCRP, OracleCRPLearnErrors, _fp_mode, simulate,  # noqa: E402
                               softmax_floor)
from oracle.rng_tape import LegacyRandom, Tape, TapeSource  # noqa: E402


def lite(model):
    ids = np.fromiter(model.cells_per_cluster.keys(), dtype=np.int64)
    sizes = np.fromiter(model.cells_per_cluster.values(), dtype=np.int64)
    return dict(assignment=np.array(model.assignment, dtype=np.int64), ids=ids, sizes=sizes,
                theta=np.array(model.parameters[ids], dtype=np.float32))


def same(a, b):
    return all(np.array_equal(a[k], b[k]) for k in ('assignment', 'ids', 'sizes', 'theta'))


def explain_gibbs(o_pre, o_tape_pos, otape, got, data, o):
    """re-run the oracle sweep from its pre-state with logging; find the first differing visit"""
    recs = otape.records
    perm = recs[o_tape_pos][1].astype(np.int64)
    assignment = o_pre['assignment'].copy()
    cpc = dict(zip(o_pre['ids'].tolist(), o_pre['sizes'].tolist()))
    o.assignment = assignment
    o.cells_per_cluster = cpc
    pos = o_tape_pos + 1
    with _fp_mode():
        fresh = o.score_new_cluster()
        for t, cell in enumerate(perm):
            was = assignment[cell]
            if cpc[was] == 1:
                del cpc[was]
            else:
                cpc[was] -= 1
            ids = np.fromiter(cpc.keys(), dtype=int)
            sizes = np.fromiter(cpc.values(), dtype=int)
            lp = np.append(o.score_existing(cell, ids), fresh[cell])
            w = softmax_floor(lp)
            u = recs[pos][1][0]
            pos += 1
            cdf = np.cumsum(w)
            cdf /= cdf[-1]
            idx = int(cdf.searchsorted(u, side='right'))
            pick = -1 if idx == ids.size else int(ids[idx])
            born = False
            if pick == -1:
                born = True
                pick = 0
                while pick in cpc:
                    pick += 1
                o.parameters[pick] = np.clip(recs[pos][1], 1e-5, 1 - 1e-5).astype(np.float32)
                pos += 1
            if got['assignment'][cell] != pick:
                print(f'  first differing visit t={t} cell={cell} was={was} oracle pick={pick} '
                      f'(born={born}) cuda={got["assignment"][cell]} u={u!r}')
                order = np.argsort(-lp)[:6]
                print('  L =', ids.size, ' position of was:', np.flatnonzero(ids == was))
                for j in order:
                    nm = 'new' if j == ids.size else f'id {ids[j]} n={sizes[j]}'
                    print(f'    pos {j:4d} {nm:18s} lp={lp[j]!r} w={w[j]!r} cdf={cdf[j]!r}')
                lo = cdf[idx - 1] if idx else 0.0
                print(f'  picked interval [{lo!r}, {cdf[idx]!r}]  margin={min(u - lo, cdf[idx] - u):.3e}')
                return
            assignment[cell] = pick
            cpc[pick] = cpc.get(pick, 0) + 1
    print('  no differing visit found (difference is in the bookkeeping)')
    print('  oracle list', list(cpc.items())[:40])
    print('  cuda   list', list(zip(got['ids'].tolist(), got['sizes'].tolist()))[:40])


def run(name, data, learning, kwargs, moves, steps, tape_arrays, init_assign):
    import torch  # noqa: F401
    from bnpc_b200.rng import Tape as PTape
    from bnpc_b200.rng import TapeRandom
    import libs.CRP as crp
    import libs.CRP_learning_errors as crple
    otape = Tape.from_arrays(*tape_arrays)
    ornd = LegacyRandom(source=TapeSource(otape))
    ocls = OracleCRPLearnErrors if learning else OracleCRP
    o = ocls(data.copy(), rnd=ornd, **kwargs)
    o.init(assign=init_assign)
    prnd = TapeRandom(PTape(*tape_arrays))
    m = (crple.CRP_errors_learning if learning else crp.CRP)(data, rnd=prnd, **kwargs)
    m.init(assign=init_assign)
    if not same(lite(o), lite(m)):
        print(f'{name}: init differs')
        return False
    for s in range(steps):
        u = ornd.random()
        assert prnd.random() == u
        if u < moves['sm_prob']:
            ro = o.update_assignments_split_merge(moves['sm_ratios'], moves['sm_steps'])
            rm = m.update_assignments_split_merge(moves['sm_ratios'], moves['sm_steps'])
            what = f'split-merge {ro} vs {rm}'
            ok = (list(ro[0]), ro[1]) == (list(rm[0]), rm[1])
        else:
            pre = lite(o)
            pos0 = otape.pos
            o.update_assignments_Gibbs()
            m.update_assignments_Gibbs()
            what = f'gibbs {m.sweep_stats}'
            ok = True
        a, b = lite(o), lite(m)
        if not ok or not same(a, b) or otape.pos != prnd.tape.pos:
            print(f'{name}: step {s + 1} {what}: MISMATCH (tape pos {otape.pos} vs {prnd.tape.pos})')
            for k in ('ids', 'sizes', 'assignment', 'theta'):
                print(f'   {k}: equal={np.array_equal(a[k], b[k])}')
            if what.startswith('gibbs'):
                explain_gibbs(pre, pos0, otape, b, data, ocls(data.copy(), rnd=ornd, **kwargs) if False else o)
            return False
        u = ornd.random()
        assert prnd.random() == u
        if u < moves['dpa_prob']:
            o.update_DP_alpha()
            m.update_DP_alpha()
        ro, rm = o.update_parameters(), m.update_parameters()
        if (int(ro[0]), int(ro[1])) != (int(rm[0]), int(rm[1])) or not same(lite(o), lite(m)):
            print(f'{name}: step {s + 1} after {what}: update_parameters differs {ro} vs {rm}')
            return False
        if learning:
            u = ornd.random()
            assert prnd.random() == u
            if u < moves['error_prob']:
                ro, rm = o.update_error_rates(), m.update_error_rates()
                if (o.FN, o.FP) != (m.FN, m.FP):
                    print(f'{name}: step {s + 1}: error rates differ')
                    return False
    print(f'{name}: {steps} steps identical')
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--golden', nargs='*', default=None)
    ap.add_argument('--smoke', action='store_true')
    args = ap.parse_args()
    names = golden_names() if args.golden is None else args.golden
    for name in names:
        g = Golden(name)
        run(name, g.data, g.meta['learning'], g.meta['kwargs'], g.meta['moves'], g.meta['steps'],
            g.tape_arrays, g.meta['init_assign'] if g.meta['init'] == 'assign' else None)
    if args.smoke:
        from oracle.crp_oracle import DEFAULT_MOVES, do_step
        data, z = simulate(400, 96, k_true=4, miss=0.1, seed=3)
        kwargs = dict(DP_alpha=[-1, -1], param_beta=[0.25, 0.25], FP_mean=0.01, FP_sd=0.01,
                      FN_mean=0.2, FN_sd=0.1)
        moves = dict(DEFAULT_MOVES, sm_prob=0.5)
        tape = Tape()
        rnd = LegacyRandom(record=tape)
        np.random.seed(1)
        o = OracleCRPLearnErrors(data.copy(), rnd=rnd, **kwargs)
        o.init()
        for _ in range(4):
            do_step(o, rnd, moves, True)
        run('smoke', data, True, kwargs, moves, 4, tape.to_arrays(), None)


if __name__ == '__main__':
    main()
