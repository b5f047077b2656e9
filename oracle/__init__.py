"""TEST INFRASTRUCTURE ONLY.

CPU restatement (torch fp32 + numpy) of the two-tower DSSM hot path of the
reference project.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package; the product package ``recommendsystemproject_b200`` never does.
"""
