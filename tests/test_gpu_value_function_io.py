"""
On-disk formats of ValueFunction (reference src/mdp.py:909-1036: column 0 `action`, then one column per state label; csv, csv.gzip,
parquet) against files WRITTEN BY THE UNMODIFIED REFERENCE (tests/golden/io/, made by tests/golden/make_golden.py io):
  * the reference's files load to exactly what the reference's own loaders return (pandas' default csv float parser is not
    round-trip exact, so "what the reference returns" is the contract, not the saved doubles);
  * our `save` / `save_parquet` write the same file contents as the reference did;
  * save -> load round trips; and the one value-function artefact the reference ships (the MDP solution of the non-wrap olfactory
    model, 20231113_182429_value_function.csv) loads through `load_from_file` to the known-answer arrays.
"""
import gzip
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu
IO = os.path.join(GOLDEN, 'io')


@pytest.fixture(scope='module')
def grid_model():
    from pomdp_pbvi_exploration_b200 import Model
    m = load_golden('model_grid4x4_noloop')
    fx = load_golden('io_grid4x4_noloop')
    S, A, O = m['rto'].shape[0], m['rto'].shape[1], m['rto'].shape[2]
    model = Model(states=[str(x) for x in fx['state_labels']], actions=A, observations=O, transitions=m['transition_table'],
                  rewards=m['reward_table'], observation_table=m['obs_table'], start_probabilities=m['start'])
    return model, fx


def test_reference_files_load_like_the_reference(grid_model):
    from pomdp_pbvi_exploration_b200 import ValueFunction
    model, fx = grid_model
    for name in ('ref_saved_grid4x4_noloop.csv', 'ref_saved_grid4x4_noloop_gz.csv.gzip'):
        vf = ValueFunction.load_from_file(os.path.join(IO, name), model)
        rows, actions = vf.numpy()
        assert np.array_equal(rows, fx['ref_loaded_csv_alphas']) and np.array_equal(actions, fx['ref_loaded_csv_actions'])
    vf = ValueFunction.load_from_parquet(os.path.join(IO, 'ref_saved_grid4x4_noloop.parquet'), model)
    rows, actions = vf.numpy()
    assert np.array_equal(rows, fx['ref_loaded_parquet_alphas']) and np.array_equal(actions, fx['ref_loaded_parquet_actions'])
    assert np.array_equal(rows, fx['alphas'])                       # parquet is exact


def test_save_writes_the_reference_format_and_round_trips(grid_model, tmp_path):
    import pandas as pd
    from pomdp_pbvi_exploration_b200 import ValueFunction
    model, fx = grid_model
    vf = ValueFunction(model, fx['alphas'], fx['actions'])
    out = str(tmp_path / 'vf')
    vf.save(path=out, file_name='mine')                              # '.csv' appended like the reference
    vf.save(path=out, file_name='mine_gz.csv', compress=True)        # '.gzip' appended, gzip-compressed
    vf.save_parquet(path=out, file_name='mine')
    assert sorted(os.listdir(out)) == ['mine.csv', 'mine.parquet', 'mine_gz.csv.gzip']
    with open(os.path.join(out, 'mine.csv')) as f, open(os.path.join(IO, 'ref_saved_grid4x4_noloop.csv')) as g:
        assert f.read() == g.read()                                  # byte-identical text to the reference's own file
    with gzip.open(os.path.join(out, 'mine_gz.csv.gzip'), 'rt') as f, open(os.path.join(IO, 'ref_saved_grid4x4_noloop.csv')) as g:
        assert f.read() == g.read()
    ours, theirs = pd.read_parquet(os.path.join(out, 'mine.parquet')), pd.read_parquet(os.path.join(IO, 'ref_saved_grid4x4_noloop.parquet'))
    assert list(ours.columns) == list(theirs.columns) and np.array_equal(ours.to_numpy(), theirs.to_numpy())
    # round trips through our own loaders: parquet exact, csv as lossy as the reference's
    back = ValueFunction.load_from_parquet(os.path.join(out, 'mine.parquet'), model)
    assert np.array_equal(back.numpy()[0], fx['alphas']) and np.array_equal(back.actions, fx['actions'])
    for name in ('mine.csv', 'mine_gz.csv.gzip'):
        back = ValueFunction.load_from_file(os.path.join(out, name), model)
        assert np.array_equal(back.numpy()[0], fx['ref_loaded_csv_alphas']) and np.array_equal(back.actions, fx['actions'])
    # default file names carry the reference's timestamp pattern
    vf.save(path=out)
    vf.save_parquet(path=out)
    names = sorted(os.listdir(out))
    assert any(n.endswith('_value_function.csv') for n in names) and any(n.endswith('_value_function.parquet') for n in names)


def test_reference_artefact_loads_through_load_from_file():
    """`Experiments/Olfactory Navigation/ValueFunctions/20231113_182429_value_function.csv` (gzip copy): 5 alpha rows over the 22021
    states of the non-wrap olfactory model; the same arrays pin VI_Solver (test_value_iteration_known_answer_olfactory_nowrap)."""
    from pomdp_pbvi_exploration_b200 import ValueFunction
    from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model
    model = olfactory_wrap_model(wrap=False)
    kat = load_golden('olf_nowrap_vi_kat')
    vf = ValueFunction.load_from_file(os.path.join(IO, 'ref_20231113_182429_value_function.csv.gzip'), model)
    rows, actions = vf.numpy()
    assert rows.shape == (5, 22021)
    assert np.array_equal(rows, kat['kat_alpha']) and np.array_equal(actions, kat['kat_action'])
    # the state labels of the recipe are the artefact's column names, so a save reproduces its header
    import pandas as pd
    cols = pd.read_csv(os.path.join(IO, 'ref_20231113_182429_value_function.csv.gzip'), compression='gzip', nrows=0).columns
    assert list(cols) == ['action', *model.state_labels]
