"""ctypes mirrors of the POD structs in include/floria_b200.h (shared by the product binding and the test oracle)."""
import ctypes as C

import numpy as np

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


class FbParams(C.Structure):
    _fields_ = [
        ("epsilon", C.c_double),
        ("div_factor", C.c_double),
        ("prob_cutoff_ln", C.c_double),
        ("max_number_solns", C.c_uint32),
        ("max_ploidy", C.c_uint32),
        ("num_iter_optimize", C.c_uint32),
        ("ploidy_sensitivity", C.c_uint32),
        ("stopping_heuristic", C.c_uint32),
        ("order_model", C.c_uint32),
        ("block_length", C.c_uint32),
        ("reassign_short", C.c_uint32),
        ("phred_lut", f32p),
    ]


class FbFrags(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint64),
        ("nnz", C.c_uint64),
        ("row_ptr", u64p),
        ("first", u32p),
        ("last", u32p),
        ("pos", u32p),
        ("allele", u8p),
        ("qual", u8p),
    ]


class FbBlockResults(C.Structure):
    _fields_ = [
        ("n_blocks", C.c_uint64),
        ("max_ploidy", C.c_uint32),
        ("_pad", C.c_uint32),
        ("best_ploidy", u32p),
        ("ploidies_run", u32p),
        ("mec_vector", f64p),
        ("expected_errors", f64p),
        ("read_ptr", u64p),
        ("read_ids", u32p),
        ("hap", u8p),
        ("cells_sweep", C.c_uint64),
        ("cells_hist", C.c_uint64),
        ("cells_beam", C.c_uint64),
        ("block_cells", u64p),
    ]


class FbParts(C.Structure):
    _fields_ = [
        ("n_parts", C.c_uint64),
        ("part_ptr", u64p),
        ("read_ids", u32p),
        ("range_lo", u32p),
        ("range_hi", u32p),
    ]


class FbBlockPhase(C.Structure):
    _fields_ = [
        ("beam_score", C.c_double),
        ("opt_score", C.c_double),
        ("n_rounds", C.c_uint32),
        ("ploidy", C.c_uint32),
        ("cells_sweep", C.c_uint64),
        ("cells_hist", C.c_uint64),
        ("cells_beam", C.c_uint64),
    ]


class FbTimings(C.Structure):
    _fields_ = [
        ("upload_ms", C.c_float),
        ("pack_ms", C.c_float),
        ("beam_ms", C.c_float),
        ("sweep_ms", C.c_float),
        ("hist_ms", C.c_float),
        ("mec_ms", C.c_float),
        ("select_ms", C.c_float),
        ("total_ms", C.c_float),
        ("download_ms", C.c_float),
        ("n_launches", C.c_uint64),
        ("n_sweep_launches", C.c_uint64),
        ("n_hist_launches", C.c_uint64),
        ("n_beam_launches", C.c_uint64),
        ("sweep_cells", C.c_uint64),
        ("hist_cells", C.c_uint64),
    ]


def ptr(a, typ):
    """Pointer of ctypes type `typ` into a C-contiguous numpy array (or None)."""
    if a is None:
        return C.cast(None, typ)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(typ)


def default_params(**kw):
    """Reference defaults: parse_cmd_line.rs:34-43,160 and constants.rs:3-6."""
    p = FbParams()
    p.epsilon = 0.04
    p.div_factor = 0.25
    p.prob_cutoff_ln = float(np.log(0.01))
    p.max_number_solns = 10
    p.max_ploidy = 5
    p.num_iter_optimize = 20
    p.ploidy_sensitivity = 2
    p.stopping_heuristic = 1
    p.order_model = 0
    p.block_length = 10000
    p.reassign_short = 0
    p.phred_lut = C.cast(None, f32p)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class BlockResults:
    """Python copy of fb_block_results (numpy arrays), so the C allocation can be freed right away."""

    def __init__(self, r):
        n = int(r.n_blocks)
        mp = int(r.max_ploidy)
        self.n_blocks = n
        self.max_ploidy = mp
        self.best_ploidy = np.ctypeslib.as_array(r.best_ploidy, (max(n, 1),))[:n].copy()
        self.ploidies_run = np.ctypeslib.as_array(r.ploidies_run, (max(n, 1),))[:n].copy()
        self.mec_vector = np.ctypeslib.as_array(r.mec_vector, (max(n * mp, 1),))[: n * mp].copy().reshape(n, mp)
        self.expected_errors = (
            np.ctypeslib.as_array(r.expected_errors, (max(n * mp, 1),))[: n * mp].copy().reshape(n, mp)
        )
        self.read_ptr = np.ctypeslib.as_array(r.read_ptr, (n + 1,)).copy()
        tot = int(self.read_ptr[n])
        self.read_ids = np.ctypeslib.as_array(r.read_ids, (max(tot, 1),))[:tot].copy()
        self.hap = np.ctypeslib.as_array(r.hap, (max(tot, 1),))[:tot].copy()
        self.cells_sweep = int(r.cells_sweep)
        self.cells_hist = int(r.cells_hist)
        self.cells_beam = int(r.cells_beam)
        self.block_cells = np.ctypeslib.as_array(r.block_cells, (max(n, 1),))[:n].copy()

    @property
    def cells(self):
        return self.cells_sweep + self.cells_hist + self.cells_beam


class Parts:
    def __init__(self, r):
        n = int(r.n_parts)
        self.n_parts = n
        self.part_ptr = np.ctypeslib.as_array(r.part_ptr, (n + 1,)).copy()
        tot = int(self.part_ptr[n])
        self.read_ids = np.ctypeslib.as_array(r.read_ids, (max(tot, 1),))[:tot].copy()
        self.range_lo = np.ctypeslib.as_array(r.range_lo, (max(n, 1),))[:n].copy()
        self.range_hi = np.ctypeslib.as_array(r.range_hi, (max(n, 1),))[:n].copy()


class FbReaderOptions(C.Structure):
    _fields_ = [("mapq_cutoff", C.c_uint32), ("use_supp_aln", C.c_uint32), ("supp_aln_dist_cutoff", C.c_int64)]


class FbFragSet(C.Structure):
    _fields_ = [
        ("frags", FbFrags),
        ("n_snps", C.c_uint64),
        ("snp_to_genome_pos", u64p),
        ("n_records", C.c_uint64),
        ("n_passed", C.c_uint64),
        ("n_without_snps", C.c_uint64),
        ("read_len_p66", C.c_uint32),
        ("_pad", C.c_uint32),
        ("contig", C.c_char * 256),
    ]
