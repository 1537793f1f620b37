"""CPU oracle for the FSNet training-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fsnet_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and there only as the checker or the timed CPU baseline.

Parity is PINNED: ``tests/golden/*.npz`` hold outputs of the reference itself
(``/root/reference`` imported read-only by ``tests/golden/make_golden.py`` in the build
container) and ``tests/test_oracle_golden.py`` checks this restatement against them.
"""
