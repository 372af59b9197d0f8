"""rkmh_b200 -- B200-native MinHash read classification (the rkmh classify/stream/filter/hash hot path).

Only what the path needs lives here: csrc/ (sm_100a kernels + the C ABI), host/ (the C++ rkmh command line),
and a thin Python mirror of the call sequence for tests and bench.py.
"""
from .api import (FILTER_TABLE_SIZE, HIT_DTYPE, STREAM_TABLE_SIZE, RkmhError, RkmhGpu, filter_pass,  # noqa: F401
                  bind_host_thread_to_gpu, pack_sequences, stream_line)
