"""Stub of the private `abltk` package (test infrastructure only).

The reference imports abltk for logging/paths/geo. This stub
lets the UNMODIFIED reference under /root/reference/src be imported in the build
container to generate golden vectors. It is never imported by bldfm_b200.
"""
