"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the OMEGA-4 analysis hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / CPU baseline -- never as the thing shipped
or measured as the GPU number.  The product (``audio-analyzer-omega_b200``) never imports
this package and fails loudly when its CUDA library is missing.
"""
