"""TESTS-ONLY shim of the 15 `diffusers==0.11.1` symbols the FateZero reference imports.

This is oracle infrastructure (see oracle/README.md): it exists so the UNMODIFIED reference package under
/root/reference can be imported in the build container to pin the CPU restatement (oracle/fz_oracle.py) and to
generate the golden vectors under tests/golden/.  It is a restatement of the published diffusers-0.11.1
behaviour, not a copy of it; nothing in the product path imports it.
"""
__version__ = "0.11.1+fz-shim"
