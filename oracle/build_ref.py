#!/usr/bin/env python3
"""Build the REAL reference (LLNL/axom, /root/reference) hot path into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (axom_b200/) may load the
library this script produces; it is the checker, never the thing measured or shipped.

What it does (no cmake, no reference build system):
  1. writes the handful of headers the reference's cmake would generate
     (axom/config.hpp, axom/mint/config.hpp from their .in templates, the
     axom/export/*.h stubs and the axom/<component>.hpp umbrella headers) into
     oracle/_ref/include/ -- outputs only, the reference tree is read where it lies;
  2. compiles the few core/slic/mint/slam .cpp files the BVH + SignedDistance
     path links against, straight from /root/reference/src, plus our own driver
     oracle/ref_driver.cpp (which calls spin::BVH<.,SEQ_EXEC> and
     quest::SignedDistance<3,SEQ_EXEC> through the reference's public API);
  3. links oracle/_ref/libaxom_ref.so.

Flags follow the reference's own Release configuration (x86-64 baseline, -O3
-DNDEBUG, no -march=native => no FMA contraction), see SURVEY.md section 8(c).
If /root/reference is absent (the GPU box) the script is a no-op: the prebuilt
.so travels with the repo snapshot.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AXOM_REFERENCE_DIR", "/root/reference")
SRC = os.path.join(REF, "src")
OUT = os.path.join(HERE, "_ref")
INC = os.path.join(OUT, "include")
OBJ = os.path.join(OUT, "obj")
LIB = os.path.join(OUT, "libaxom_ref.so")

DEFINES = {
    "AXOM_VERSION_MAJOR": "0",
    "AXOM_VERSION_MINOR": "11",
    "AXOM_VERSION_PATCH": "0",
    "AXOM_VERSION_FULL": "v0.11.0",
    "AXOM_SRC_DIR_NATIVE": SRC,
    "AXOM_BIN_DIR_NATIVE": OUT,
    "BLT_CXX_STD": "c++14",
    "AXOM_DEPRECATED_TYPES_N": "2",
    "AXOM_MSVC_PRAGMAS": "",
    "AXOM_GIT_SHA": "dc40842",
}
# what the survey's cmake configuration turned on (no RAJA / Umpire / MPI / OpenMP,
# 32-bit IndexType) -- SURVEY.md 8(c)
ENABLED = {
    "AXOM_USE_CLI11", "AXOM_USE_FMT",  # sparsehash off: MapCollection falls back to std::unordered_map
    "AXOM_USE_MINT", "AXOM_USE_PRIMAL", "AXOM_USE_QUEST", "AXOM_USE_SLAM",
    "AXOM_USE_SLIC", "AXOM_USE_SPIN", "AXOM_DEPRECATED_TYPES_N",
}
ENABLED01 = {"AXOM_FMT_EXCEPTIONS": 1, "AXOM_FMT_HEADER_ONLY": 1}

COMPONENT_SOURCES = {
    "core": ["utilities/Annotations.cpp", "utilities/FileUtilities.cpp",
             "utilities/StringUtilities.cpp", "utilities/System.cpp",
             "utilities/Utilities.cpp", "numerics/polynomial_solvers.cpp",
             "Path.cpp", "Types.cpp"],
    "slic": ["core/Logger.cpp", "core/LogStream.cpp", "interface/slic.cpp",
             "internal/stacktrace.cpp", "streams/GenericOutputStream.cpp"],
    "mint": ["mesh/internal/MeshHelpers.cpp", "mesh/blueprint.cpp",
             "mesh/CurvilinearMesh.cpp", "mesh/FieldData.cpp", "mesh/Mesh.cpp",
             "mesh/MeshCoordinates.cpp", "mesh/ParticleMesh.cpp",
             "mesh/RectilinearMesh.cpp", "mesh/StructuredMesh.cpp",
             "mesh/UniformMesh.cpp", "fem/FiniteElement.cpp"],
    "slam": ["BitSet.cpp", "OrderedSet.cpp"],
    "quest": ["SignedDistance.cpp", "interface/signed_distance.cpp", "interface/internal/QuestHelpers.cpp",
              "readers/STLReader.cpp", "readers/ProEReader.cpp", "MeshTester.cpp"],
}

# NOTE: the image exports CXX=/opt/gcc/bin/g++, a wrapper without libgomp; use the system g++.
CXX = os.environ.get("AXB_HOST_CXX", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")
CXXFLAGS = ["-O3", "-DNDEBUG", "-std=c++14", "-fPIC", "-w", "-ffp-contract=off", "-fopenmp"]


def configure(template, dest):
    """mini configure_file(): #cmakedefine / #cmakedefine01 / @VAR@."""
    out = []
    for line in open(template):
        m = re.match(r"\s*#\s*cmakedefine01\s+(\w+)", line)
        if m:
            out.append("#define %s %d\n" % (m.group(1), ENABLED01.get(m.group(1), 0)))
            continue
        m = re.match(r"\s*#\s*cmakedefine\s+(\w+)(.*)", line)
        if m:
            name, rest = m.group(1), m.group(2)
            if name in ENABLED:
                rest = re.sub(r"@(\w+)@", lambda mm: DEFINES.get(mm.group(1), ""), rest)
                out.append("#define %s%s\n" % (name, rest.rstrip()))
            else:
                out.append("/* #undef %s */\n" % name)
            continue
        out.append(re.sub(r"@(\w+)@", lambda mm: DEFINES.get(mm.group(1), ""), line))
    os.makedirs(os.path.dirname(dest), exist_ok=True)
    with open(dest, "w") as f:
        f.writelines(out)


def component_headers(comp):
    """the set(<comp>_headers ...) list of the component's CMakeLists.txt, as
    axom_write_unified_header() would see it (detail/ and internal/ excluded)."""
    txt = open(os.path.join(SRC, "axom", comp, "CMakeLists.txt")).read()
    m = re.search(r"set\(\s*%s_headers(.*?)\)" % comp, txt, re.S)
    hdrs = []
    for tok in m.group(1).split("\n"):
        tok = tok.split("#")[0].strip()
        if tok.endswith(".hpp") or tok.endswith(".h"):
            if "/detail/" in "/" + tok or "/internal/" in "/" + tok:
                continue
            if os.path.exists(os.path.join(SRC, "axom", comp, tok)):
                hdrs.append(tok)
    return hdrs


def write_generated_headers():
    configure(os.path.join(SRC, "axom", "config.hpp.in"), os.path.join(INC, "axom", "config.hpp"))
    configure(os.path.join(SRC, "axom", "mint", "core", "config.hpp.in"),
              os.path.join(INC, "axom", "mint", "config.hpp"))
    os.makedirs(os.path.join(INC, "axom", "export"), exist_ok=True)
    for c in ("slic", "mint", "slam", "mir", "sidre"):
        C = c.upper()
        with open(os.path.join(INC, "axom", "export", c + ".h"), "w") as f:
            f.write("#ifndef AXOM_%s_EXPORT_H\n#define AXOM_%s_EXPORT_H\n"
                    "#define AXOM_%s_EXPORT\n#define AXOM_%s_NO_EXPORT\n#endif\n" % (C, C, C, C))
    for comp in ("core", "slic", "primal", "mint", "slam", "spin"):
        hdrs = component_headers(comp)
        if comp == "mint":
            hdrs = ["config.hpp"] + hdrs
        with open(os.path.join(INC, "axom", comp + ".hpp"), "w") as f:
            f.write("#ifndef AXOM_UNIFIED_%s_HPP\n#define AXOM_UNIFIED_%s_HPP\n" % (comp.upper(), comp.upper()))
            f.write('#include "axom/config.hpp"\n')
            for h in hdrs:
                f.write('#include "axom/%s/%s"\n' % (comp, h))
            f.write("#endif\n")
    # axom/fmt.hpp itself is the reference's thirdparty/axom/fmt.hpp, found through -isystem
    # About.cpp is a configured source in the reference's build (core/CMakeLists.txt:16-19)
    os.makedirs(os.path.join(OUT, "gen"), exist_ok=True)
    configure(os.path.join(SRC, "axom", "core", "utilities", "About.cpp.in"), os.path.join(OUT, "gen", "About.cpp"))


def compile_one(src, obj, extra=()):
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return
    cmd = [CXX] + CXXFLAGS + list(extra) + [
        "-I" + INC, "-I" + SRC, "-isystem", os.path.join(SRC, "thirdparty"),
        "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stderr[-6000:] + "\n")
        raise SystemExit("reference compile failed: " + src)


def main():
    if not os.path.isdir(SRC):
        print("[build_ref] %s not present: keeping prebuilt oracle/_ref (if any)" % SRC)
        return 0
    os.makedirs(OBJ, exist_ok=True)
    write_generated_headers()
    jobs = []
    for comp, files in COMPONENT_SOURCES.items():
        for f in files:
            src = os.path.join(SRC, "axom", comp, f)
            obj = os.path.join(OBJ, comp + "_" + f.replace("/", "_").replace(".cpp", ".o"))
            jobs.append((src, obj))
    jobs.append((os.path.join(OUT, "gen", "About.cpp"), os.path.join(OBJ, "core_About.o")))
    driver = os.path.join(HERE, "ref_driver.cpp")
    jobs.append((driver, os.path.join(OBJ, "ref_driver.o")))
    # quest::MarchingCubes needs Conduit, which is not in this image: the reference's two sources + our driver are compiled
    # against the Node MOCK in oracle/conduit_stub/ (a named tree of external arrays; no algorithm), reference files unmodified
    mc_flags = ("-DAXOM_USE_CONDUIT", "-I" + os.path.join(HERE, "conduit_stub"))
    mc_jobs = [(os.path.join(SRC, "axom", "quest", "MarchingCubes.cpp"), os.path.join(OBJ, "quest_MarchingCubes.o"), mc_flags),
               (os.path.join(SRC, "axom", "quest", "detail", "MarchingCubesSingleDomain.cpp"),
                os.path.join(OBJ, "quest_detail_MarchingCubesSingleDomain.o"), mc_flags),
               (os.path.join(HERE, "ref_mc_driver.cpp"), os.path.join(OBJ, "ref_mc_driver.o"), mc_flags)]
    jobs = [j + ((),) for j in jobs] + mc_jobs
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(lambda j: compile_one(*j), jobs))
    jobs = [j[:2] for j in jobs]
    objs = [o for _, o in jobs]
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [CXX, "-shared", "-fopenmp", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr[-6000:])
            raise SystemExit("reference link failed")
    print("[build_ref] built", LIB)
    return 0


if __name__ == "__main__":
    sys.exit(main())
