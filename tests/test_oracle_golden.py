"""CPU: the oracle restatement replays the tapes recorded from the REFERENCE and must
land on the reference's states bit for bit (tests/golden/*.npz, made by make_golden.py)."""
import numpy as np
import pytest

from helpers import Golden, assert_state, golden_names
from oracle.crp_oracle import OracleCRP, OracleCRPLearnErrors, do_step, snapshot
from oracle.rng_tape import LegacyRandom, TapeSource


@pytest.mark.parametrize('name', golden_names())
def test_oracle_replays_reference_tape(name):
    g = Golden(name)
    tape = g.oracle_tape()
    rnd = LegacyRandom(source=TapeSource(tape))
    cls = OracleCRPLearnErrors if g.meta['learning'] else OracleCRP
    model = cls(g.data.copy(), rnd=rnd, **g.meta['kwargs'])
    model.init(assign=g.meta['init_assign'] if g.meta['init'] == 'assign' else None)
    assert tape.pos == g.tape_pos[1]
    assert_state(snapshot(model), g.state(0), f'{name} init')
    for s in range(g.meta['steps']):
        log = do_step(model, rnd, g.meta['moves'], g.meta['learning'])
        assert log == g.steplog[s], f'{name} step {s + 1}: {log} vs {g.steplog[s]}'
        assert tape.pos == g.tape_pos[s + 2], f'{name} step {s + 1}: tape position'
        assert_state(snapshot(model), g.state(s + 1), f'{name} step {s + 1}')
    assert tape.exhausted()


def test_fixtures_cover_all_moves():
    seen = set()
    for name in golden_names():
        for log in Golden(name).steplog:
            seen.add(log['move'])
            if log['sm'] and log['sm'][0]:
                seen.add(log['move'] + '_accepted')
            if log['errors']:
                seen.add('errors')
            if log['alpha']:
                seen.add('alpha')
    assert {'gibbs', 'split', 'merge', 'split_accepted', 'merge_accepted', 'errors', 'alpha'} <= seen
