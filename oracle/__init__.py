"""TEST INFRASTRUCTURE — CPU oracle for the OptCuts hot path.

Nothing under ``oracle/`` is on the product path: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.

* ``oracle/_ref/``      the UNMODIFIED reference, compiled from /root/reference by ``oracle/Makefile``
                        (git-ignored; travels to the GPU box as a prebuilt)
* ``oracle/refapi.py``  ctypes binding of ``_ref/liboptcuts_ref.so`` (the real reference classes)
* ``oracle/port/``      our plain-C restatement of the path + ``oracle/portapi.py`` binding
* ``oracle/state_io.py`` reader for the probe's binary state dumps
"""
