"""B200-native LTE channel-decoding engine (turbo decoder, rate dematching, sub-block
deinterleaving) behind the OpenAirInterface entry points.

The product is the CUDA library ``openair4g_b200/lib/liboai_turbo_b200.so`` (C ABI in
``include/oai_turbo_b200.h``).  ``openair4g_b200.capi`` is a thin ctypes mirror of that
ABI with the reference's function names.  There is no CPU compute path in this package:
importing ``capi`` raises if the CUDA library has not been built.
"""
__version__ = "0.1"
