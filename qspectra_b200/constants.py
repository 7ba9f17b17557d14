"""Unit-conversion constants (values must equal the reference's exactly:
reference qspectra/constants.py:8-18)."""
import math

#: fs -> linear cm^-1
CM_FS_LINEAR = 2.99792458e-5
#: fs -> angular cm^-1
CM_FS = math.pi * 2 * CM_FS_LINEAR
#: Kelvin -> angular cm^-1 (Boltzmann constant folded in)
CM_K = 0.69503476
#: sigma / FWHM of a Gaussian
GAUSSIAN_SD_FWHM = 1.0 / (2 * math.sqrt(2 * math.log(2)))
