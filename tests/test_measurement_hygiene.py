"""The DRAM-traffic figure bench.py puts into `roofline.traffic` comes from an ncu capture committed under profiles/.  It is only
meaningful for the kernel sources it was captured from: bench.py drops it (traffic = null, with a note) when the hash of the
kernel sources differs, and this test makes the mismatch loud at commit time."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_committed_traffic_capture_matches_the_kernel_sources():
    import bench
    prof = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
    assert prof["kernel_source_hash"] == bench.kernel_source_hash(), (
        "profiles/dram_traffic.json was captured from other kernel sources: re-run "
        "`tools/gpu_round.sh <tag> ncu_launches ncu_full` on the GPU box and `python tools/summarize_profiles.py <tag> 8192`")
    assert prof["ls_bp_kernel_dram_bytes_per_problem_pass"] > 0
    src = prof["source"].split(" ")[0]
    assert os.path.exists(os.path.join(ROOT, src)), src   # the raw ncu metrics the figure was taken from are committed too
