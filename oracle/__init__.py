"""ORACLE -- test infrastructure only.

CPU ground truth for the AGBNP1 / GaussVol hot path.  Two arms:
  * oracle.reflib  -- ctypes binding of oracle/_ref/libagbnp_ref.so: the reference's own UNMODIFIED sources
                      (gaussvol.cpp, AGBNPForce.cpp, AGBNPUtils.cpp, ReferenceAGBNPKernels.cpp) compiled against
                      oracle/shim (see oracle/Makefile, oracle/ref_driver.cpp).  kind = "reference".
  * oracle.portlib -- ctypes binding of oracle/libagbnp_oracle.so: oracle/agbnp_oracle.c, a plain-C restatement
                      (each function cites the reference file:line it follows) which also carries the
                      cutoff-aware variant the Reference platform lacks.  kind = "port".
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (openmm_agbnp_plugin_b200) never does.
"""
