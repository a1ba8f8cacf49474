"""fesom2-accelerate_b200 -- B200-native fct_ale tracer limiter behind the fesom2-accelerate C ABI.

The directory name carries the reference's hyphen, so import it with
    importlib.import_module("fesom2-accelerate_b200")
(the repo root on sys.path).  Nothing here imports torch or the oracle.
"""
