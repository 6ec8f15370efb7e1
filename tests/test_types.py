"""Result containers: replay, cost/latency accessors and the JSON layout.  No GPU."""
import numpy as np
from conftest import golden_cases, load_golden

from da4ml_b200.types import CombLogic, Pipeline, pipeline_from_arrays


def test_pipeline_reproduces_kernel_and_roundtrips(tmp_path):
    for name, meta in golden_cases().items():
        extra, stages = load_golden(name)
        for st in stages:
            n_out = len(st['out_idxs'])
            st['shape'] = (len(st['inp_shifts']), n_out)
            st['carry_size'] = meta['kwargs'].get('carry_size', -1)
            st['adder_size'] = meta['kwargs'].get('adder_size', -1)
        pipe = pipeline_from_arrays(stages)
        want = extra['kernel'].copy()
        if 'qint' in extra:  # inputs whose interval is {0, 0} are dropped by the solver (state_opr.cc:92-97)
            want[(extra['qint'][:, 0] == 0) & (extra['qint'][:, 1] == 0)] = 0
        assert np.array_equal(pipe.kernel, want), name
        assert pipe.n_adders == meta['n_adders']
        x = np.random.default_rng(0).integers(-8, 8, size=(5, pipe.shape[0])).astype(np.float64)
        assert np.array_equal(pipe(x), x @ want.astype(np.float64))
        p = tmp_path / f'{name}.json'
        pipe.save(p)
        again = Pipeline.load(p)
        assert again == pipe
        assert isinstance(again.solutions[0], CombLogic)
        assert again.cost == pipe.cost and again.latency == pipe.latency
