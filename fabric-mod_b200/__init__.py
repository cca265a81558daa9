"""fabric-mod_b200 -- B200-native batch ECDSA-P256 verification behind Fabric's bccsp.BCCSP.Verify.

The product is the C-ABI library ``lib/libfabgpu_ecdsa.so`` (sources in ``csrc/``, ABI in
``include/fabgpu_ecdsa.h``).  This Python package is the thin host-side harness used by tests and bench.py:

  * ``binding``  -- ctypes view of the C ABI (one method per exported function);
  * ``bccsp``    -- mirror of the reference provider interface for this path
                    (bccsp.BCCSP.Verify / KeyImport, sw.CSP error behaviour; reference bccsp/bccsp.go:90-134,
                    bccsp/sw/impl.go:247-270) plus msp identity.Verify (reference msp/identities.go:169-196).

There is no CPU fallback in here: if the CUDA library cannot be loaded or no device exists, calls raise.
The directory name contains a hyphen, so import it with ``importlib.import_module("fabric-mod_b200")``.
"""
from . import binding, bccsp  # noqa: F401

__all__ = ["binding", "bccsp"]  # `sharding` imports torch; load it explicitly where needed
