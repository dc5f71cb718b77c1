"""Synthetic weights and inputs for the oracle side (TEST INFRASTRUCTURE): re-exports the package's pure data generators
(segmentation-networks-benchmark_b200/synth.py) so that golden vectors, oracle checks and the device path all draw from
the same seeds."""
import snb_b200  # noqa: F401  (registers the package under its importable name)
from snb_b200.synth import *  # noqa: F401,F403
from snb_b200.synth import _bias, _he  # noqa: F401
