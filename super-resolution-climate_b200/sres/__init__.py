"""Host-side mirror of the reference's `sres` package for the RCAN hot path:
same module paths, names, argument meaning and error behaviour, with the arithmetic on
libsres_b200.so.  Put `super-resolution-climate_b200/` on sys.path and `import sres` as before."""
