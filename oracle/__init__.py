"""CPU oracle for the per-frame estimation path of wear_mocap_ape 1.2.3.

TEST INFRASTRUCTURE ONLY.  A numpy / torch-CPU restatement of the reference's algorithm (every function
cites the reference file:line it follows).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or as
the timed CPU baseline - never from the product package ``arm_pose_estimation_b200``.

Pinning: the reference has no tests or golden vectors of its own, so the oracle is pinned
against outputs of the reference itself, run in the build container by ``tests/golden/make_golden.py``
(unmodified ``wear_mocap_ape`` classes, imported from /root/reference) and committed under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them on every CPU test run.

Third-party arithmetic on the path that is not under /root/reference: ``torch.nn.LSTM`` + inter-layer
dropout + ``torch.nn.Linear`` (torch, unpinned in the reference's setup.cfg:24; 2.11.0 here), restated in
``oracle/lstm.py`` from the published cell equations and checked against ``torch.nn.LSTM`` itself, and
``numpy.linalg.eigh`` (numpy 2.3.5 here), which the oracle calls exactly as the reference does.
"""
