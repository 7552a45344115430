"""Address targets of sample/observe statements (CPU): same information as the reference's bytecode inspection
(pyprob/state.py:31-84): '<instruction pointer>__<function chain>__<assignment target>' with the target a variable name,
`base[int]` for subscript stores with a constant or local integer index, 'return', or '?'."""
from pyprob_b200 import state


def _probe():
    return state._extract_address(2)     # the frame that called _probe, like sample() -> _addresses() -> _extract_address


def _model():
    out = {}
    a = _probe()
    out['a'] = a
    xs = [None, None, None]
    xs[0] = _probe()
    i = 2
    xs[i] = _probe()
    key = 'name'
    table = {}
    table[key] = _probe()
    out['xs'], out['table'] = xs, table
    out['bare'] = [_probe()]
    return out


def _returns():
    return _probe()


def test_assignment_targets(monkeypatch):
    monkeypatch.setattr(state, '_root_function_name', '_model')
    got = _model()
    assert got['a'].endswith('___model__a') and got['a'].split('__')[0].isdigit()
    assert got['xs'][0].endswith('___model__xs[0]')          # constant index  (state.py:66-72)
    assert got['xs'][2].endswith('___model__xs[2]')          # local integer index (:73-76)
    assert got['table']['name'].endswith('___model__?')      # non-integer index -> no target (:80-81)
    assert got['bare'][0].endswith('___model__?')
    monkeypatch.setattr(state, '_root_function_name', '_returns')
    assert _returns().endswith('___returns__return')
    # distinct statements get distinct instruction pointers
    ips = {got['a'].split('__')[0], got['xs'][0].split('__')[0], got['xs'][2].split('__')[0]}
    assert len(ips) == 3
