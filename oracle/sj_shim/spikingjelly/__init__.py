"""Shim of the un-vendored third-party ``spikingjelly==0.0.0.0.14`` (TEST INFRASTRUCTURE).

Only the surface the reference touches; semantics restated in
``oracle/plif.py``.  Put ``oracle/sj_shim`` on ``sys.path`` to let the reference's model
code import unmodified in this container.  PARITY UNPINNED (see oracle/__init__.py).
"""
