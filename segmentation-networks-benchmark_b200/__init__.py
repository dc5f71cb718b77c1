"""B200-native tiled-segmentation hot path of BloodAxe/segmentation-networks-benchmark.

The directory name carries the reference's name and is not a Python identifier; import it as `snb_b200`
(the repo-root shim `snb_b200.py` registers this directory as that package).  `snb_b200.lib.*` mirrors the
reference's `lib.*` modules for the path inria_submit.py drives: tiles.ImageSlicer, models.UNet16/UNet11,
losses.BCEWithLogitsLossAndSmoothJaccard, metrics.JaccardScore/PixelAccuracy.  All compute happens in
libsnb_b200.so (hand-written sm_100a CUDA behind the C ABI of include/snb_b200.h); nothing falls back to the CPU.
"""
__version__ = "0.1.0"
