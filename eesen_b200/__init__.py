"""eesen_b200 -- B200-native (sm_100a) CTC-training hot path for Eesen.

The product is the C-ABI library ``eesen_b200/lib/libeesen_b200.so`` (hand-written CUDA kernels in
``csrc/`` + the C++ host mirror of Eesen's Net/Layer/Ctc API in ``host/``) and the
``train-ctc-parallel`` driver.  The Python modules are plumbing for tests and ``bench.py``:
``binding`` (ctypes over include/eesen_b200.h), ``kaldi_io`` (on-disk formats), ``synth``
(seeded synthetic workloads of the BASELINE.json configs).
"""
__version__ = "0.1.0"
