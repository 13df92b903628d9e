"""rx_tools_b200 — Blackwell-native I/Q DSP engine behind the rx_tools hot path.

Host-side mirror of the reference's operator interface for the two paths SURVEY.md §8 names:

* ``rx_tools_b200.fm``    — rx_fm: stream-callback DSP body + ``full_demod()``
* ``rx_tools_b200.power`` — rx_power: ``scanner()`` per-hop window x ``fix_fft()`` x power

Both call the C-ABI library ``librxb200.so`` (include/rxb200.h, built from ``csrc/``) through
ctypes; there is no CPU fallback — if the library or a CUDA device is missing the calls raise.
"""
__all__ = ["synth"]
