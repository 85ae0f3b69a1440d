"""Host logic of bench.py that needs no GPU: the HF schedule choice (child-process probes are replaced by canned outputs)."""
import argparse
import json

import pytest

import bench


def _args(hf_lanes="auto"):
    return argparse.Namespace(hf_lanes=hf_lanes, workload="synth8k", contexts=24)


def _fake(monkeypatch, identical, speeds):
    """Canned child outputs keyed by the schedule: value + digest ("ref" when identical to the default kernel's)."""
    canned = {}
    for n, v in speeds.items():
        canned[n] = {"value": v, "sha256": "ref" if n == "0" or identical.get(n) else "other"}
    monkeypatch.setenv("JXLB_BENCH_FAKE_PROBE", json.dumps(canned))
    monkeypatch.delenv("JXLB_HF_LANES", raising=False)


def test_candidate_must_be_identical_and_faster(monkeypatch):
    _fake(monkeypatch, {"16": True, "64": True, "128": True}, {"0": 3000.0, "16": 3300.0, "64": 3600.0, "128": 3900.0})
    n, rep = bench.choose_hf_schedule(_args(), 0)
    assert n == 128 and rep["chosen"] == 128 and rep["probe_mp_s"]["0"] == 3000.0
    _fake(monkeypatch, {"16": True, "64": True, "128": False}, {"0": 3000.0, "16": 3700.0, "64": 3600.0, "128": 9999.0})
    n, rep = bench.choose_hf_schedule(_args(), 0)
    assert n == 16 and rep["identical_to_default_kernel"]["128"] is False   # differs from the default kernel: never used
    _fake(monkeypatch, {"16": True, "64": True, "128": True}, {"0": 3000.0, "16": 3010.0, "64": 3050.0, "128": 2900.0})
    assert bench.choose_hf_schedule(_args(), 0)[0] == 0           # within 3 %: keep the default
    _fake(monkeypatch, {"16": True}, {"0": 3000.0, "16": 3500.0})  # the 64 / 128 children "crash": only they are dropped
    n, rep = bench.choose_hf_schedule(_args(), 0)
    assert n == 16 and str(rep["identical_to_default_kernel"]["64"]).startswith("failed")


def test_probe_failure_falls_back_to_the_default(monkeypatch):
    monkeypatch.setenv("JXLB_BENCH_FAKE_PROBE", json.dumps({}))     # the parity child "fails"
    monkeypatch.delenv("JXLB_HF_LANES", raising=False)
    n, rep = bench.choose_hf_schedule(_args(), 0)
    assert n == 0 and "fallback" in rep


@pytest.mark.parametrize("flag,env,want", [("64", None, 64), ("auto", "128", 128), ("0", None, 0), ("auto", "50", 64), ("16", None, 16)])
def test_explicit_choice_wins(monkeypatch, flag, env, want):
    monkeypatch.delenv("JXLB_BENCH_FAKE_PROBE", raising=False)
    if env is None:
        monkeypatch.delenv("JXLB_HF_LANES", raising=False)
    else:
        monkeypatch.setenv("JXLB_HF_LANES", env)
    n, rep = bench.choose_hf_schedule(_args(flag), 0)
    assert n == want and rep["mode"] == "explicit"
