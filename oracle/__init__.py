"""CPU oracle for the DGVCC density-supervision hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dgvcc_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker or the reported CPU baseline -- never as the thing shipped.

Each module restates one reference file's arithmetic on the CPU (torch-CPU /
numpy / scipy, the same third-party primitives the reference itself calls):

* ``bl_oracle``   <- /root/reference/losses/bl.py
* ``dmap_oracle`` <- /root/reference/utils/dmap_gen.py
* ``isw_oracle``  <- /root/reference/models/ISW/instance_whitening.py

Pinning: the reference ships no tests or golden vectors,
so the pins are outputs of the unmodified reference files executed in the
authoring container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
``tests/test_oracle_*.py`` check every restatement against those fixtures.
"""
