__version__ = "0.32.0.dev0+orv_b200_shim"
