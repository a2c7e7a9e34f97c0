"""Drop-in module name of the reference (`#[pymodule] fn vfind`, src/lib.rs:322-327)."""
from vfind_b200 import find_variants  # noqa: F401

__all__ = ["find_variants"]
