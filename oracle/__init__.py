"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference Seeker forward.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker or the timed CPU baseline.
``tcow_b200`` never imports this package.
"""
