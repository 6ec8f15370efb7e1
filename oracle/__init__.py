"""CPU checkers for the CMVM path -- TEST INFRASTRUCTURE ONLY.

  oracle/ref.py   -> oracle/_ref/libcmvm_ref.so   the reference's own TUs compiled in place (strongest checker)
  oracle/port.py  -> oracle/libcmvm_oracle.so     from-scratch restatement (cmvm_oracle.cc), travels everywhere

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (da4ml_b200) never does.
"""

STAGE_KEYS = ['inp_shifts', 'out_idxs', 'out_shifts', 'out_negs', 'ops_i', 'ops_f']


def best():
    """The strongest available checker module and its kind ('reference' or 'port')."""
    from . import port, ref

    if ref.available():
        return ref, 'reference'
    return port, 'port'


def oracle_solve(kernel, **kw):
    mod, kind = best()
    return mod.solve(kernel, **kw), kind
