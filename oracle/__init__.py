"""oracle/ -- TEST INFRASTRUCTURE ONLY (see the headers of the modules in here)."""
import os

# libgomp reads OMP_STACKSIZE once, when it is first loaded -- by liboracle.so or by libhaccref*.so, whichever comes
# first.  The compiled reference keeps 4*VMAX floats on each OpenMP worker's stack (RCBForceTree.cxx:940; 16 MB in
# the _vmax build), so the value must be in the environment before either library is opened.
os.environ.setdefault("OMP_STACKSIZE", "64M")
