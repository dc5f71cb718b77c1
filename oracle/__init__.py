"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Restates, in numpy / CPU PyTorch, the reference algorithms of the hot path
(BloodAxe/segmentation-networks-benchmark: lib/tiles.py, lib/augmentations.py:452-511, lib/common.py:59-76,
lib/losses.py:31-75, lib/metrics.py, lib/train_utils.py:92-131, lib/models/unet11.py, lib/models/unet16.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import it, and only
as the checker or the reported CPU baseline; nothing under segmentation-networks-benchmark_b200/ imports it.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against the reference itself, imported from /root/reference in the build container by oracle/make_golden.py;
the resulting input/output vectors are committed under tests/golden/ and tests/test_oracle_golden.py checks
every oracle function against them on any machine.
"""
