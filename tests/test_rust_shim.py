"""SURVEY §8(f).1 — the Rust shim crate `bft-lib-gpu/` cannot be compiled here (no cargo/rustc in the image), so
its FFI surface is checked textually against include/lbft.h: `#[repr(C)]` structs field by field (order, width,
signedness, pointer-ness), every `extern "C"` function (name, arity, argument kinds), the flag constants, and the
link flags build.rs emits."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "lbft.h")).read()
RUST = open(os.path.join(ROOT, "bft-lib-gpu", "src", "lib.rs")).read()

C2KIND = {"uint32_t": "u32", "int32_t": "i32", "uint64_t": "u64", "int64_t": "i64", "double": "f64", "uint8_t": "u8",
          "size_t": "usize", "int": "c_int", "char": "c_char", "void": "u8"}


def strip_comments(text):
    return re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", text, flags=re.S))


def c_struct_fields(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), strip_comments(HEADER), re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(const\s+)?(\w+)\s*(\*)?\s*(.*)$", decl)
        const, ctype, ptr, names = m.groups()
        for nm in names.split(","):
            nm = nm.strip()
            arr = re.match(r"(\w+)\[(\d+)\]", nm)
            if arr:
                out.append((arr.group(1), "[%s; %s]" % (C2KIND[ctype], arr.group(2))))
            else:
                kind = C2KIND[ctype]
                out.append((nm, ("*const " if const else "*mut ") + kind if ptr else kind))
    return out


def rust_struct_fields(name):
    body = re.search(r"pub struct %s \{(.*?)\n    \}" % name, strip_comments(RUST), re.S).group(1)
    return [(m.group(1), m.group(2).strip()) for m in re.finditer(r"pub (\w+): ([^,\n]+),", body)]


def test_repr_c_structs_match_the_header():
    pairs = {"lbft_config": "LbftConfig", "lbft_commit": "LbftCommit", "lbft_round_switch": "LbftRoundSwitch",
             "lbft_instance_counters": "LbftInstanceCounters"}
    for cname, rname in pairs.items():
        c, r = c_struct_fields(cname), rust_struct_fields(rname)
        c = [("lambda" if n == "lambda" else n, t) for n, t in c]
        assert c == r, "%s vs %s:\n%s\n%s" % (cname, rname, c, r)
        assert "#[repr(C)]" in RUST.split("pub struct %s" % rname)[0][-200:]


def c_functions():
    out = {}
    for m in re.finditer(r"\n(?:const\s+)?(\w+)\s*(\*)?\s*(lbft_\w+)\s*\(([^)]*)\)\s*;", strip_comments(HEADER)):
        ret, retptr, name, args = m.groups()
        kinds = []
        for a in [a.strip() for a in args.split(",")]:
            if a in ("void", ""):
                continue
            am = re.match(r"(const\s+)?(\w+)\s*(\*+)?\s*(\w+)?$", a)
            const, ctype, stars, _ = am.groups()
            base = C2KIND.get(ctype, ctype)
            kinds.append(("ptr:" + base) if stars else base)
        out[name] = (kinds, (("ptr:" if retptr else "") + C2KIND.get(ret, ret)))
    return out


def rust_functions():
    block = re.search(r'extern "C" \{(.*?)\n    \}', strip_comments(RUST), re.S).group(1)
    out = {}
    for m in re.finditer(r"pub fn (lbft_\w+)\(([^)]*)\)(?:\s*->\s*([^;]+))?;", block):
        name, args, ret = m.groups()
        kinds = []
        for a in [a.strip() for a in args.split(",") if a.strip()]:
            t = a.split(":", 1)[1].strip()
            pm = re.match(r"\*(?:const|mut)\s+(.*)$", t)
            kinds.append("ptr:" + pm.group(1) if pm else t)
        out[name] = (kinds, (ret or "void").strip())
    return out


RUST2C = {"LbftConfig": "lbft_config", "LbftSim": "lbft_sim", "LbftCommit": "lbft_commit", "LbftRoundSwitch": "lbft_round_switch",
          "LbftInstanceCounters": "lbft_instance_counters", "*mut LbftSim": "ptr:lbft_sim"}


def norm(kind):
    if kind.startswith("ptr:"):
        inner = kind[4:]
        if inner.startswith("*mut "):   # pointer to pointer
            return "ptr:" + RUST2C.get(inner[5:], inner[5:])
        return "ptr:" + RUST2C.get(inner, inner)
    return kind


def test_extern_functions_match_the_header():
    c, r = c_functions(), rust_functions()
    assert set(r) <= set(c), sorted(set(r) - set(c))
    # the shim binds everything its GpuSimulator needs, including the batch, async and bulk-log entry points
    for needed in ("lbft_create", "lbft_run", "lbft_run_async", "lbft_wait", "lbft_run_until", "lbft_commit_counts", "lbft_last_states",
                   "lbft_commit_log", "lbft_commit_logs", "lbft_round_switches", "lbft_destroy", "lbft_last_error"):
        assert needed in r, needed
    for name, (rk, rret) in r.items():
        ck, cret = c[name]
        assert [norm(k) for k in rk] == [norm(k).replace("ptr:void", "ptr:u8") for k in ck], (name, rk, ck)
        assert norm(rret.replace("*const ", "ptr:")) == cret.replace("int", "c_int") if cret == "int" else True, (name, rret, cret)


def test_flag_constants_and_link_flags():
    for cname in ("LBFT_FLAG_ROUND_SWITCHES", "LBFT_FLAG_RESUMABLE", "LBFT_FLAG_TRUE_DATA_SYNC"):
        cval = int(re.search(r"#define %s (\d+)u" % cname, HEADER).group(1))
        rval = int(re.search(r"pub const %s: u32 = (\d+);" % cname, RUST).group(1))
        assert cval == rval, cname
    build = open(os.path.join(ROOT, "bft-lib-gpu", "build.rs")).read()
    assert "cargo:rustc-link-lib=dylib=lbft_b200" in build and "cargo:rustc-link-search=native=" in build
    cargo = open(os.path.join(ROOT, "bft-lib-gpu", "Cargo.toml")).read()
    assert 'build = "build.rs"' in cargo and "bft-lib" in cargo and "librabft-v2" in cargo
    # the golden values of the reference's own integration test are what the crate's test asserts
    t = open(os.path.join(ROOT, "bft-lib-gpu", "tests", "simulated_run.rs")).read()
    for golden in ("11134312813757838303", "12785928431398617538", "4890275890002623733"):
        assert golden in t
