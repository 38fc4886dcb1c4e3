"""Writes the build-configuration header utility/Version.hpp for the oracle/_ref build
(the reference's CMake would configure it from core/CMake/Spirit_Version.hpp.in).
TEST INFRASTRUCTURE -- only used by oracle/Makefile."""
fields = [
    ("int", "version_major", 2), ("int", "version_minor", 2), ("int", "version_patch", 0),
    ("str", "version", "2.2.0"), ("str", "version_revision", "oracle"), ("str", "version_full", "2.2.0 (oracle)"),
    ("str", "compiler", "GNU"), ("str", "compiler_version", "13"), ("str", "compiler_full", "GNU (13)"),
    ("str", "scalartype", "double"), ("str", "pinning", "OFF"), ("str", "defects", "OFF"),
    ("str", "cuda", "OFF"), ("str", "openmp", "ON"), ("str", "threads", "OFF"), ("str", "fftw", "OFF"),
]
print("#pragma once\n#include <string>\nnamespace Utility\n{")
for kind, name, val in fields:
    if kind == "int":
        print(f"const int {name} = {val};")
    else:
        print(f'const std::string {name} = "{val}";')
print("}")
