"""CPU: host-side pieces of bench.py that the GPU runs depend on -- the sub-record watchdog (a hung collective in an
optional sub-record must never cost the headline line) and the argument defaults the driver relies on."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, timeout=60):
    return subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_watchdog_prints_the_headline_line_and_exits_zero():
    r = _run("import bench, time\n"
             "d = bench.SubRecordWatchdog(0, {'metric': 'm', 'value': 1.5})\n"
             "d.arm('other_configs', 1)\n"
             "time.sleep(30)\n"
             "print('NOT REACHED')\n")
    assert r.returncode == 0 and "NOT REACHED" not in r.stdout
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line == {"metric": "m", "value": 1.5, "sub_records_timed_out": "other_configs"}
    assert "exceeded its time budget" in r.stderr


def test_watchdog_other_ranks_exit_silently_and_disarm_works():
    r = _run("import bench, time\n"
             "d = bench.SubRecordWatchdog(3, None)\n"
             "d.arm('train', 1)\n"
             "time.sleep(30)\n")
    assert r.returncode == 0 and r.stdout.strip() == ""
    r = _run("import bench, time\n"
             "d = bench.SubRecordWatchdog(0, {'value': 2})\n"
             "d.arm('train', 1)\n"
             "d.disarm()\n"
             "time.sleep(3)\n"
             "print('alive')\n")
    assert r.returncode == 0 and r.stdout.strip() == "alive"


def test_driver_facing_defaults():
    r = _run("import sys, bench\n"
             "sys.argv = ['bench.py']\n"
             "a = bench.parse_args()\n"
             "print(a.gpus, a.batch, a.views, a.new_tokens, a.layers, a.impl, a.subrecord_timeout)\n")
    assert r.returncode == 0, r.stderr
    assert r.stdout.split() == ["1", "128", "6", "256", "32", "b200", "300.0"]


def test_algorithmic_flop_counts_match_the_survey():
    """configs[1] = 16.95 TFLOP per inference (6 views, L = 831, 256 new tokens); the fine-tune step of
    configs[4] counts forward + 2 x forward of everything that is trained (179.3 TFLOP at 4 x 981 tokens per GPU)."""
    r = _run("import bench\n"
             "from mm_or_b200.config import LlavaConfig\n"
             "from mm_or_b200.train.bench_step import step_flops\n"
             "print(round(bench.algorithmic_flops_per_inference(6, 831, 256) / 1e12, 2))\n"
             "print(round(step_flops(LlavaConfig(num_hidden_layers=32), 4, 6, 981) / 1e12, 1))\n")
    assert r.returncode == 0, r.stderr
    assert r.stdout.split() == ["16.95", "179.3"]
