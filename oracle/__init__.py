"""CPU oracle for the RIFE-4.6 hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only
as the checker or the reported CPU baseline - never as the thing measured or shipped.
The product path (``comfyui-frame-interpolation_b200``) never imports this package and
fails loudly when its CUDA library is missing.
"""
