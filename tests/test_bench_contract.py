"""bench.py's reference arm runs on CPU: check the JSON contract (keys, units, reference-arm extras).  No GPU."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--size', '16', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['unit'] == 'matrices/s' and line['higher_is_better'] is True
    assert line['metric'] == 'cmvm_solve_throughput_16x16_int8' and 'workload' in line['config']
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert line['vs_baseline'] is None


def test_default_metric_is_the_baseline_metric():
    sys.path.insert(0, str(ROOT))
    import bench

    assert bench.metric_name(256, 8) == bench.METRIC == 'cmvm_solve_throughput_256x256_int8'
    spec = json.loads((ROOT / 'BASELINE.json').read_text())
    assert '256' in spec['metric'] and 'int8' in spec['metric']
