"""Golden texts for da4ml_b200.emit from the reference's OWN modules: ``trace/pipeline.py::to_pipeline`` (retiming off),
``codegen/rtl/verilog``, ``codegen/rtl/vhdl``, ``codegen/hls/hls_codegen.py`` and ``rtl_model.binder_gen``, run on the adder
graphs of the committed golden solves.

Container only: imports /root/reference/src/da4ml as a package with its native module replaced by a stand-in
(tests/ref_trace.py; none of the emitters calls into it).  Writes tests/golden/codegen.json.gz.
"""
import gzip
import importlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE.parent))

CASES = ['c1_8x8_int4_default', 'int_16x16_int8_default', 'pytest_8_b4_harddc2_add1', 'single_16x12_hetero_wmc', 'int_12x20_int6_harddc1', 'int_17x5_int8_mcpdc',
         'pytest_4_b2_mc_wmc']  # fmt: skip


def _neg_zero():
    rng = np.random.default_rng(7)
    W = rng.integers(-32, 32, size=(6, 7)).astype(np.float32)
    W[:, 2] = 0  # a dead output (out_idx -1)
    W[:, 3] = -np.abs(W[:, 3]) - 1  # outputs the solver negates
    W[:, 5] = -W[:, 4]  # ... and one that is the negation of another column
    W[1, :] = 0  # a dead input
    return W, dict(adder_size=1, carry_size=4)  # (adders of several latency units: retiming has something to move)


def _neg_frac():
    rng = np.random.default_rng(11)
    W = (rng.integers(-64, 64, size=(9, 6)) * rng.choice([0, 1], size=(9, 6), p=[0.4, 0.6])).astype(np.float32) * np.float32(0.125)
    W[:, 0] = -np.abs(W[:, 0])
    q = [(-8.0, 7.5, 0.5), (0.0, 3.0, 1.0), (-1.0, 0.0, 1.0), (0.0, 0.0, 1.0), (-128.0, 127.0, 1.0), (0.0, 15.75, 0.25), (-4.0, 4.0, 2.0), (-0.5, 0.5, 0.5), (1.0, 9.0, 1.0)]
    return W, dict(qintervals=q, adder_size=2, carry_size=3, hard_dc=1)


def _dead_input():
    rng = np.random.default_rng(13)
    W = rng.integers(-32, 32, size=(7, 6)).astype(np.float32)
    W[2, :] = 0  # an input nothing reads
    q = [(-128.0, 127.0, 1.0)] * 7
    q[4] = (0.0, 0.0, 1.0)  # ... and one whose interval is a point: dropped by the solver (state_opr.cc:92-97), a constant to the reference tracer
    return W, dict(adder_size=2, carry_size=4, qintervals=q)


CUSTOM = {'custom_neg_zero_6x7': _neg_zero, 'custom_neg_frac_9x6': _neg_frac, 'custom_dead_input_7x6': _dead_input}


def stage_lists(sol):
    """CombLogic -> plain lists (what the tests compare the pipelined stages with)."""
    return dict(shape=list(sol.shape), inp_shifts=list(sol.inp_shifts), out_idxs=list(sol.out_idxs), out_shifts=list(sol.out_shifts), out_negs=[bool(v) for v in sol.out_negs],
                ops=[[o.id0, o.id1, o.opcode, o.data, list(o.qint), o.latency, o.cost] for o in sol.ops])  # fmt: skip


def main():
    import ref_trace
    from conftest import golden_cases, load_golden

    from da4ml_b200.types import pipeline_from_arrays

    from oracle import port

    T, _ = ref_trace.load(None, port.get_lsb_loc, port.iceil_log2, port.cost_add)  # (the tracer behind retime_pipeline calls cost_add)
    P = importlib.import_module('da4ml.trace.pipeline')
    V = importlib.import_module('da4ml.codegen.rtl.verilog')
    VH = importlib.import_module('da4ml.codegen.rtl.vhdl')
    H = importlib.import_module('da4ml.codegen.hls.hls_codegen')
    R = importlib.import_module('da4ml.codegen.rtl.rtl_model')
    out = {}
    for name in CASES + list(CUSTOM):
        if name in CUSTOM:  # solved here by the reference's object code (oracle/_ref); the arrays travel inside the json
            from oracle import ref

            W, kw = CUSTOM[name]()
            stages = [dict(st) for st in ref.solve(W, **kw)]
            meta = {'kwargs': kw}
        else:
            meta = golden_cases()[name]
            _, stages = load_golden(name)
        for st in stages:
            st['shape'] = (len(st['inp_shifts']), len(st['out_idxs']))
            st['carry_size'] = meta['kwargs'].get('carry_size', -1)
            st['adder_size'] = meta['kwargs'].get('adder_size', -1)
        pipe = pipeline_from_arrays(stages, types_module=T)  # the reference's own NamedTuples
        assert type(pipe).__module__ == 'da4ml.types'
        rec = {'stages': []}
        if name in CUSTOM:
            rec['arrays'] = [{k: np.asarray(st[k]).tolist() for k in ('inp_shifts', 'out_idxs', 'out_shifts', 'out_negs', 'ops_i')} | {'ops_f_bits': np.asarray(st['ops_f'], np.float32).view(np.uint32).tolist(),
                                                                                                                                   'adder_size': st['adder_size'], 'carry_size': st['carry_size']} for st in stages]
        for i, sol in enumerate(pipe.solutions):
            s = {
                'verilog': V.comb_logic_gen(sol, f'm{i}', print_latency=False, timescale=None),
                'verilog_lat': V.comb_logic_gen(sol, f'm{i}', print_latency=True, timescale='`timescale 1ns/1ps'),
                'vhdl': VH.comb_logic_gen(sol, f'm{i}', print_latency=(i == 1)),
                'verilog_io': V.generate_io_wrapper(sol, f'm{i}', False),
                'vhdl_io': VH.generate_io_wrapper(sol, f'm{i}', False),
                'binder': R.binder_gen(sol, f'm{i}_wrapper'),
                'hls': {fl: list(H.hls_logic_and_bridge_gen(sol, f'f{i}', fl, pragmas=['#pragma HLS INLINE'] if fl == 'vitis' else None, print_latency=(fl == 'hlslib'),
                                                            namespace='ns' if fl == 'oneapi' else '', n_base_indent=1 if fl == 'oneapi' else 0))
                        for fl in ('vitis', 'hlslib', 'oneapi')},
                'to_pipeline': {},
            }  # fmt: skip
            lat = max(sol.out_latency) if sol.out_latency else 0.0
            small = len(sol.ops) < 120
            for cut in sorted({2.0, max(lat / 2.0, 1.0)} | ({1.0, 3.5} if small else set())) if (i == 0 or small) else [2.0]:
                try:
                    csol = P.to_pipeline(sol, cut, retiming=False)
                except Exception as e:  # (a cutoff that leaves a stage empty is a KeyError in the reference)
                    s['to_pipeline'][repr(cut)] = {'error': type(e).__name__}
                    continue
                rec_c = {'stages': [stage_lists(c) for c in csol.solutions]}
                try:  # the same split balanced by the reference's retime_pipeline (re-traces through its symbolic tracer)
                    rec_c['retimed'] = [stage_lists(c) for c in P.to_pipeline(sol, cut, retiming=True, verbose=False).solutions]
                except Exception as e:
                    rec_c['retimed'] = {'error': type(e).__name__}
                for key, fn in {
                    'verilog': lambda: V.pipeline_logic_gen(csol, 'pm', register_layers=1),
                    'verilog_r3': lambda: V.pipeline_logic_gen(csol, 'pm', print_latency=True, register_layers=3)['pm'],
                    'vhdl': lambda: VH.pipeline_logic_gen(csol, 'pm', register_layers=2),
                    'verilog_io': lambda: V.generate_io_wrapper(csol, 'pm', True),
                    'vhdl_io': lambda: VH.generate_io_wrapper(csol, 'pm', True),
                    'binder': lambda: R.binder_gen(csol, 'pm_wrapper', 1, 2),
                }.items():
                    try:
                        rec_c[key] = fn()
                    except Exception as e:  # (inputs that arrive late keep their original index in a later stage: the reference's emitters reject that)
                        rec_c[key] = {'error': type(e).__name__}
                s['to_pipeline'][repr(cut)] = rec_c
            rec['stages'].append(s)
        # the solver's own two-stage result as a register pipeline (what RTLModel does with latency_cutoff <= 0)
        try:
            rt = P.retime_pipeline(pipe, verbose=False)
            rec['pipeline_retimed'] = [stage_lists(c) for c in rt.solutions]
            if name in CUSTOM or len(pipe.solutions[0].ops) < 120:  # the text the reference emits for the retimed pipeline
                rec['pipeline_retimed_text'] = {
                    'verilog': V.pipeline_logic_gen(rt, 'rt'),
                    'vhdl': VH.pipeline_logic_gen(rt, 'rt'),
                    'hls': [list(H.hls_logic_and_bridge_gen(c, f'rt{k}', 'vitis')) for k, c in enumerate(rt.solutions)],
                    'verilog_io': V.generate_io_wrapper(rt, 'rt', True),
                }
        except Exception as e:
            rec['pipeline_retimed'] = {'error': type(e).__name__}
        rec['pipeline'] = {
            'verilog': V.pipeline_logic_gen(pipe, 'top'),
            'vhdl': VH.pipeline_logic_gen(pipe, 'top'),
            'verilog_io': V.generate_io_wrapper(pipe, 'top', True),
            'vhdl_io': VH.generate_io_wrapper(pipe, 'top', True),
            'binder': R.binder_gen(pipe, 'top_wrapper'),
        }
        out[name] = rec
    raw = json.dumps(out, separators=(',', ':')).encode()
    with gzip.GzipFile(HERE / 'codegen.json.gz', 'wb', mtime=0) as f:
        f.write(raw)
    print(len(out), 'cases,', len(raw) // 1024, 'KiB of text,', (HERE / 'codegen.json.gz').stat().st_size // 1024, 'KiB compressed')


if __name__ == '__main__':
    main()
