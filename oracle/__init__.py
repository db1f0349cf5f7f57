"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference hot path.

Nothing under ``po2_quantization_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it,
and only as the checker / the CPU arm, never as the product path.
"""
