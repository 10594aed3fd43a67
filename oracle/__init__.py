"""Test infrastructure only: CPU restatement of pyHGT's HGTConv hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only
as the checker (or the timed CPU baseline), never on the CUDA product path.
"""
