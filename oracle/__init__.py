"""Test infrastructure only: CPU restatement of the reference (``lav_ref``), the script that pins it against the unmodified
reference (``pin_against_reference``) and the import shims that script needs (``refshim``).  Nothing under ``lav_b200``
imports this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference / cpu_baseline legs do."""
