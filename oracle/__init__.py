"""CPU oracles: RIFE 4.6 / 4.7 / 4.17 / 4.26 (rife46.py), FILM (film.py), Sepconv (sepconv.py), GMFSS Fortuna
(gmflow.py + gmfss.py; target only, the GPU path is not built), the op primitives (ops_ref.py).
TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only
as the checker or the reported CPU baseline - never as the thing measured or shipped.
The product path (``comfyui-frame-interpolation_b200``) never imports this package and
fails loudly when its CUDA library is missing.
"""
