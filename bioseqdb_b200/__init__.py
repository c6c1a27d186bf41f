"""bioseqdb_b200: the read-alignment hot path of unneon/bioseqdb (nuclseq_search_bwa /
nuclseq_multi_search_bwa) rebuilt for NVIDIA B200 (sm_100a): hand-written CUDA kernels behind a C ABI
(include/bioseqdb_gpu.h), with this package as the host-side mirror of the reference's adapter interface."""
from .bwa import AlignResult, BwaIndex, BwaMatch, MultiBwaIndex, bwa_opts, cigar_to_string  # noqa: F401
from .cache import BwaIndexCache, rows_digest  # noqa: F401
from .sequence import NucleotideSequence, nuclseq_from_text, nuclseq_to_text  # noqa: F401
from ._lib import BsqError, BsqOpts, build_library  # noqa: F401
