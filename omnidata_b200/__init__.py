"""omnidata_b200 — B200-native (sm_100a) DPT-Hybrid-384 dense-prediction path of EPFL-VILAB/omnidata.

Python host code above a C-ABI CUDA library (include/omnidata_b200.h).  No CPU fallback.
"""
__version__ = "0.1.0"
