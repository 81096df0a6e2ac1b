"""
Cassandra `.POMDP` file reader / writer (the reference's `load_POMDP_file`, src/pomdp.py:3383-3736).

`load_POMDP_file(path) -> (Model, PBVI_Solver)` like the reference.  The format (https://pomdp.org/code/pomdp-file-spec.html):

    discount: <float>      values: reward|cost
    states: <n> | <names...>      actions: ...      observations: ...
    start: uniform | <state> | <p_0 ... p_{S-1}>  (probabilities may follow on the next line)
    start include: <states...>    start exclude: <states...>
    T: <a> : <s> : <s'> <p>       T: <a> : <s>  <row | uniform>        T: <a>  <matrix | uniform | identity>
    O: <a> : <s'> : <o> <p>       O: <a> : <s'> <row | uniform>        O: <a>  <matrix | uniform>
    R: <a> : <s> : <s'> : <o> <r> R: <a> : <s> : <s'> <row over o>     R: <a> : <s>  <matrix s' x o>

with `*` as a wildcard for any index, names or integers for states / actions / observations, `#` comments and values
that may continue on following lines.  This is a tokenising parser written for this package (entries are applied in file
order, later ones override earlier ones); it accepts the full grammar above, which includes the three example files the
reference's line-oriented reader rejects.  `save_POMDP_file` writes a model back in the explicit
`T/O/R: ... value` form (round-trip tested).  Host-side I/O only: nothing here touches the device.
"""
from __future__ import annotations

import re
from typing import Tuple

import numpy as np

from .model import Model

_KEYWORDS = ('discount', 'values', 'states', 'actions', 'observations', 'start', 'T', 'O', 'R')


def _tokens(text: str) -> list:
    """Comment-free token stream in which every ':' is its own token and each keyword that opens a statement is tagged."""
    out = []
    for raw in text.splitlines():
        line = raw.split('#', 1)[0].strip()
        if not line:
            continue
        line = line.replace(':', ' : ')
        parts = line.split()
        head = parts[0]
        if head in _KEYWORDS and len(parts) > 1 and (parts[1] == ':' or (head == 'start' and parts[1] in ('include', 'exclude'))):
            if head == 'start' and parts[1] in ('include', 'exclude'):
                out.append(('KEY', 'start ' + parts[1]))
                parts = parts[3:] if len(parts) > 2 and parts[2] == ':' else parts[2:]
            else:
                out.append(('KEY', head))
                parts = parts[2:]
        out.extend(('TOK', p) for p in parts)
    return out


def _is_number(tok: str) -> bool:
    return re.fullmatch(r'[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?', tok) is not None


def _names(tokens: list, prefix: str) -> list:
    if len(tokens) == 1 and tokens[0].isdigit():
        return [f'{prefix}_{i}' for i in range(int(tokens[0]))]
    return list(tokens)


class _Space:
    def __init__(self, labels):
        self.labels = labels
        self.n = len(labels)
        self.index = {lab: i for i, lab in enumerate(labels)}

    def ids(self, tok: str) -> list:
        if tok == '*':
            return list(range(self.n))
        if tok in self.index:
            return [self.index[tok]]
        if tok.isdigit() and int(tok) < self.n:
            return [int(tok)]
        raise ValueError(f"unknown name '{tok}' (expected one of {self.labels[:8]}...)")


def parse_POMDP(text: str) -> dict:
    """Parses the text of a .POMDP file into dense tables: dict(discount, values, states, actions, observations,
    transitions [S,A,S], observation_table [S,A,O] (indexed by the landing state), rewards [S,A,S,O], start [S])."""
    toks = _tokens(text)
    # group the stream into statements: (keyword, [tokens until the next keyword])
    stmts = []
    for kind, val in toks:
        if kind == 'KEY':
            stmts.append([val, []])
        else:
            if not stmts:
                raise ValueError(f"token '{val}' before any statement")
            stmts[-1][1].append(val)
    head = {k: v for k, v in stmts if k in ('discount', 'values', 'states', 'actions', 'observations')}
    for need in ('discount', 'states', 'actions', 'observations'):
        if need not in head:
            raise ValueError(f"missing '{need}:' statement")
    S_, A_, O_ = _Space(_names(head['states'], 's')), _Space(_names(head['actions'], 'a')), _Space(_names(head['observations'], 'o'))
    S, A, O = S_.n, A_.n, O_.n
    T = np.zeros((S, A, S))
    Z = np.zeros((S, A, O))
    Rw = np.zeros((S, A, S, O))
    start = np.full(S, 1.0 / S)
    sign = -1.0 if head.get('values', ['reward'])[0] == 'cost' else 1.0

    def split_colon(body):
        fields, cur = [], []
        for t in body:
            if t == ':':
                fields.append(cur)
                cur = []
            else:
                cur.append(t)
        fields.append(cur)
        return fields

    for key, body in stmts:
        if key in head:
            continue
        if key == 'start':
            if body == ['uniform']:
                start = np.full(S, 1.0 / S)
            elif len(body) == 1 and not _is_number(body[0]):
                start = np.zeros(S)
                start[S_.ids(body[0])] = 1.0
            elif len(body) == 1 and S > 1:
                start = np.zeros(S)
                start[S_.ids(body[0])] = 1.0
            else:
                if len(body) != S:
                    raise ValueError(f'start: expected {S} probabilities, found {len(body)}')
                start = np.array([float(x) for x in body])
        elif key in ('start include', 'start exclude'):
            ids = sorted({i for t in body for i in S_.ids(t)})
            mask = np.zeros(S, dtype=bool)
            mask[ids] = True
            if key == 'start exclude':
                mask = ~mask
            start = mask / mask.sum()
        elif key == 'T':
            f = split_colon(body)
            acts = A_.ids(f[0][0])
            if len(f) == 1:                                   # T: a  <matrix | uniform | identity>
                vals = f[0][1:]
                mat = _matrix(vals, S, S, allow_identity=True)
                for a in acts:
                    T[:, a, :] = mat
            elif len(f) == 2:                                 # T: a : s  <row | uniform>
                rows = S_.ids(f[1][0])
                row = _row(f[1][1:], S)
                for a in acts:
                    T[rows, a, :] = row
            else:                                             # T: a : s : s' p
                rows, cols = S_.ids(f[1][0]), S_.ids(f[2][0])
                p = float(f[2][1])
                for a in acts:
                    T[np.ix_(rows, [a], cols)] = p
        elif key == 'O':
            f = split_colon(body)
            acts = A_.ids(f[0][0])
            if len(f) == 1:
                mat = _matrix(f[0][1:], S, O, allow_identity=False)
                for a in acts:
                    Z[:, a, :] = mat
            elif len(f) == 2:
                rows = S_.ids(f[1][0])
                row = _row(f[1][1:], O)
                for a in acts:
                    Z[rows, a, :] = row
            else:
                rows, cols = S_.ids(f[1][0]), O_.ids(f[2][0])
                p = float(f[2][1])
                for a in acts:
                    Z[np.ix_(rows, [a], cols)] = p
        elif key == 'R':
            f = split_colon(body)
            acts, srcs = A_.ids(f[0][0]), S_.ids(f[1][0])
            if len(f) == 2:                                   # R: a : s  <matrix s' x o>
                mat = np.array([float(x) for x in f[1][1:]]).reshape(S, O) * sign
                for a in acts:
                    for s in srcs:
                        Rw[s, a, :, :] = mat
            elif len(f) == 3:                                 # R: a : s : s'  <row over o>
                dsts = S_.ids(f[2][0])
                row = np.array([float(x) for x in f[2][1:]]) * sign
                if row.shape[0] != O:
                    raise ValueError(f'R: expected {O} values, found {row.shape[0]}')
                for a in acts:
                    Rw[np.ix_(srcs, [a], dsts, range(O))] = row
            else:                                             # R: a : s : s' : o r
                dsts, obs = S_.ids(f[2][0]), O_.ids(f[3][0])
                r = float(f[3][1]) * sign
                for a in acts:
                    Rw[np.ix_(srcs, [a], dsts, obs)] = r
    return dict(discount=float(head['discount'][0]), values=head.get('values', ['reward'])[0], states=S_.labels, actions=A_.labels,
                observations=O_.labels, transitions=T, observation_table=Z, rewards=Rw, start=start)


def _row(vals: list, n: int) -> np.ndarray:
    if vals == ['uniform']:
        return np.full(n, 1.0 / n)
    if len(vals) != n:
        raise ValueError(f'expected {n} values, found {len(vals)}')
    return np.array([float(x) for x in vals])


def _matrix(vals: list, rows: int, cols: int, allow_identity: bool) -> np.ndarray:
    if vals == ['uniform']:
        return np.full((rows, cols), 1.0 / cols)
    if vals == ['identity']:
        if not allow_identity or rows != cols:
            raise ValueError("'identity' is only valid for transition matrices")
        return np.eye(rows)
    if len(vals) != rows * cols:
        raise ValueError(f'expected a {rows}x{cols} matrix, found {len(vals)} values')
    return np.array([float(x) for x in vals]).reshape(rows, cols)


def load_POMDP_file(file_name: str) -> Tuple[Model, object]:
    """(Model, PBVI_Solver(gamma=discount)) from a .POMDP file, as the reference's loader returns them."""
    from .solver import PBVI_Solver
    with open(file_name) as f:
        spec = parse_POMDP(f.read())
    model = Model(states=spec['states'], actions=spec['actions'], observations=spec['observations'], transitions=spec['transitions'],
                  rewards=spec['rewards'], observation_table=spec['observation_table'], start_probabilities=spec['start'])
    return model, PBVI_Solver(gamma=spec['discount'])


def save_POMDP_file(model: Model, file_name: str, discount: float) -> None:
    """Writes a model with dense tables as a .POMDP file (explicit entries for every non-zero probability / reward)."""
    assert model.transition_table is not None and model.immediate_reward_table is not None, 'dense transition / reward tables are required'
    S, A, O = model.state_count, model.action_count, model.observation_count
    T, Z, Rw = model.transition_table, model.observation_table, model.immediate_reward_table
    with open(file_name, 'w') as f:
        f.write(f'discount: {discount!r}\nvalues: reward\nstates: {S}\nactions: {A}\nobservations: {O}\n\n')
        f.write('start:\n' + ' '.join(repr(float(p)) for p in model.start_probabilities) + '\n\n')
        for a in range(A):
            for s in range(S):
                for sn in np.flatnonzero(T[s, a]):
                    f.write(f'T: {a} : {s} : {sn} {float(T[s, a, sn])!r}\n')
        for a in range(A):
            for sn in range(S):
                for o in np.flatnonzero(Z[sn, a]):
                    f.write(f'O: {a} : {sn} : {o} {float(Z[sn, a, o])!r}\n')
        for s, a, sn, o in zip(*np.nonzero(Rw)):
            f.write(f'R: {a} : {s} : {sn} : {o} {float(Rw[s, a, sn, o])!r}\n')
