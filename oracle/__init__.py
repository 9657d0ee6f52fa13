"""CPU oracle for the uncalibrated-visual-servoing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``uncalibrated-visual-servoing_b200``) never imports this package
and fails loudly when its CUDA library is missing.

What is here is a plain NumPy restatement of the reference's per-frame
estimators and the IBVS step, each function citing the reference file:line it
follows (paths relative to /root/reference):

* ``filters``      experiment.py:164-297 (KF / MCKF / IMCC-KF / GMCKF=RMCKF, as
                   written, dense) + the tests/ variants (simple-form P,
                   no-skip MCKF) + the exact block-structured restatement.
* ``ibvs``         experiment.py:300-312.
* ``kinematics``   ur10_simulation.py:97-154,204-211 + experiment.py:94-114 and
                   the scene constants of SURVEY.md Appendix A.
* ``noise_model``  noise.py:81-207 restated on counter-based Philox4x32-10
                   (same distributions, different streams).
* ``stats``        results/plot_errorbar.m:39-94.
* ``closed_loop``  experiment.py:48-359 as one function on the kinematic plant.
* ``fake_robot``   duck-typed stand-ins for ``UR10Simulation`` used to drive
                   the UNMODIFIED reference (``make_golden.py``).

Pinning: the reference ships no golden vectors or tests for this path
(SURVEY.md §4).  The oracle is therefore pinned against outputs of the
reference itself run in the build container: ``oracle/make_golden.py`` imports
``/root/reference/experiment.py`` + ``noise.py`` unmodified, drives them with
``fake_robot.FakeRobot`` and writes per-frame X / P / dq / logs to
``tests/golden/*.npz``; ``tests/test_oracle_vs_golden.py`` requires the
restatement to reproduce them (<=1e-12 relative).  Philox is pinned by the
Random123 known-answer vectors.
"""
