"""The Rust shim (shim/src/core/vqb200_ffi.rs) cannot be compiled here (no cargo / rustc), so its `extern "C"` block is
checked textually against include/vqb200.h: same set of functions, same argument count, and every argument / return type
the Rust spelling of the C one.  The struct vqb_train_opts and the callback typedefs are checked field by field too.
Template: the reference's own binding, src/core/hsdlib_ffi.rs:38-66."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "vqb200.h")
FFI = os.path.join(ROOT, "shim", "src", "core", "vqb200_ffi.rs")

C2RUST = {"int": "c_int", "void": "c_void", "size_t": "usize", "float": "f32", "uint8_t": "u8", "uint16_t": "u16",
          "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "char": "c_char", "vqb_ctx": "VqbCtx", "vqb_pq": "VqbPq",
          "vqb_tsvq": "VqbTsvq", "vqb_train_opts": "VqbTrainOpts", "vqb_reseed_fn": "ReseedFn",
          "vqb_allreduce_fn": "AllreduceFn"}


def rust_type(c):
    c = " ".join(c.split())
    const = c.startswith("const ")
    if const:
        c = c[6:]
    stars = c.count("*")
    t = C2RUST[c.replace("*", "").strip()]
    for i in range(stars):
        t = ("*const " if (const and i == 0) else "*mut ") + t
    return t


def header_functions():
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"^\s*((?:const\s+)?[A-Za-z_]\w*\s*\*?)\s*(vqb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.M | re.S):
        args = " ".join(args.split())
        params = []
        if args not in ("", "void"):
            for a in args.split(","):
                params.append(rust_type(re.match(r"(.*?)([A-Za-z_]\w*)$", a.strip()).group(1)))
        out[name] = (rust_type(ret), params)
    return out


def shim_functions():
    src = re.sub(r"//.*", "", open(FFI).read())
    block = re.search(r'unsafe extern "C" \{(.*?)\n\}', src, flags=re.S).group(1)
    out = {}
    for name, args, ret in re.findall(r"pub fn (vqb_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        params = [" ".join(a.split(":", 1)[1].split()) for a in args.split(",") if a.strip()]
        out[name] = (" ".join(ret.split()) if ret else "c_void", params)
    return out


def test_every_header_function_is_bound_with_the_same_signature():
    h, r = header_functions(), shim_functions()
    assert len(h) >= 40, "header parse lost functions"
    assert sorted(h) == sorted(r), f"only in header: {sorted(set(h) - set(r))}; only in shim: {sorted(set(r) - set(h))}"
    for name in h:
        assert h[name] == r[name], f"{name}: header {h[name]} != shim {r[name]}"


def test_train_opts_layout_matches():
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    body = re.search(r"typedef struct vqb_train_opts \{(.*?)\} vqb_train_opts;", src, flags=re.S).group(1)
    c_fields = []
    for line in body.split(";"):
        line = " ".join(line.split())
        if line:
            m = re.match(r"(.*?)([A-Za-z_]\w*)$", line)
            c_fields.append((m.group(2), rust_type(m.group(1))))
    rs = re.sub(r"//.*", "", open(FFI).read())
    rbody = re.search(r"pub struct VqbTrainOpts \{(.*?)\}", rs, flags=re.S).group(1)
    r_fields = [(n, " ".join(t.split())) for n, t in re.findall(r"pub (\w+):\s*([^,]+),", rbody)]
    assert c_fields == r_fields
    assert "#[repr(C)]\npub struct VqbTrainOpts" in open(FFI).read()


def test_callback_typedefs_and_constants_match():
    hdr = open(HDR).read()
    rs = open(FFI).read()
    assert re.search(r"typedef uint64_t \(\*vqb_reseed_fn\)\(void\* user, uint32_t subspace\);", hdr)
    assert re.search(r"pub type ReseedFn = Option<unsafe extern \"C\" fn\(user: \*mut c_void, subspace: u32\) -> u64>;", rs)
    assert re.search(r"typedef int \(\*vqb_allreduce_fn\)\(void\* user, float\* buf, size_t count, void\* cuda_stream\);", hdr)
    assert re.search(r"fn\(user: \*mut c_void, buf: \*mut f32, count: usize, cuda_stream: \*mut c_void\) -> c_int", rs)
    for cname, rname in (("VQB_UPDATE_ORDERED", "VQB_UPDATE_ORDERED"), ("VQB_UPDATE_FAST", "VQB_UPDATE_FAST"),
                         ("VQB_ASSIGN_AUTO", "VQB_ASSIGN_AUTO"), ("VQB_ASSIGN_EXACT", "VQB_ASSIGN_EXACT"),
                         ("VQB_ASSIGN_TENSOR", "VQB_ASSIGN_TENSOR"), ("VQB_COMM_ID_BYTES", "VQB_COMM_ID_BYTES"),
                         ("VQB_CHEBYSHEV", "VQB_CHEBYSHEV")):
        cv = int(re.search(rf"#define {cname}\s+(\d+)", hdr).group(1))
        rv = int(re.search(rf"pub const {rname}: \w+ = (\d+);", rs).group(1))
        assert cv == rv, cname
    for code, variant in ((0, "Success"), (-1, "ErrNullPtr"), (-2, "ErrEmptyInput"), (-3, "ErrInvalidInput"),
                          (-4, "ErrUnsupportedDevice"), (-5, "ErrDimMismatch"), (-99, "Failure")):
        assert re.search(rf"{variant} = {code},", rs), variant


def test_shim_files_are_present_and_cite_the_reference():
    for rel, cite in (("build.rs", "build.rs:6-40"), ("src/pq.rs", "pq.rs:91-117"), ("src/tsvq.rs", "tsvq.rs:31-115"),
                      ("src/bq.rs", "bq.rs:94-105"), ("src/sq.rs", "sq.rs:123-151"), ("src/core/vqb200_ffi.rs", "hsdlib_ffi.rs")):
        text = open(os.path.join(ROOT, "shim", rel)).read()
        assert cite in text, (rel, cite)
