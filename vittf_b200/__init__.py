"""vittf_b200 -- B200-native (sm_100a) implementation of the vit-tf feature-volume hot path.

Python keeps the reference's call surface (infer.py / predict_ntf.py / bilateral_solver3d.py
function names and signatures); all arithmetic on the path runs in libvittf_b200.so through the C ABI
declared in include/vittf.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
