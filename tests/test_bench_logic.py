"""Host logic of bench.py that needs no GPU."""
import argparse

import bench


def test_metric_names_follow_the_workload():
    assert "8K VarDCT d1.0" in bench.METRIC["synth8k"]
    assert "d2.0" in bench.METRIC["synth8k_d2"] and "Modular" in bench.METRIC["synthmod4k"]


def test_workloads_are_seeded_and_sized():
    desc, frames, (w, h) = bench.load_workload("synth4k", 5)
    assert (w, h) == (3840, 2160) and len(frames) == 5 and frames[0] == frames[4] and frames[0] != frames[1]
    assert "EPF 2" in desc


def test_fixed_hf_schedule_and_roofline_families():
    assert bench.HF_STREAMS_PER_CTA in (8, 16, 32, 64, 128) and bench.HF_LATENCY_SCHEDULE in (8, 16, 32)
    assert set(bench.CHAIN) & set(bench.KERNELS) and set(bench.ENTROPY) <= set(bench.KERNELS)
    assert bench.algorithmic_bytes("filters_fused", 7680, 4320, 0) == 7680 * 4320 * 24


def test_gpu_local_cpus_without_a_gpu_is_empty():
    assert bench.gpu_local_cpus(0) == []
