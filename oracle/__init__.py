"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product package
(``reveal_b200``) never does; it fails loudly when its CUDA library is missing.

Two checkers live here:

* ``oracle.port``  -- ctypes wrapper over ``oracle/reveal_oracle.c`` (a plain-C
  restatement of the reference path, pinned against the next item), and
* ``oracle.ref``   -- loader for the UNMODIFIED reference extension compiled
  from ``/root/reference`` into ``oracle/_ref/`` by ``oracle/ref/Makefile``.
"""
