"""Host mirror of ``orpheum/translate_single_seq.py:7-78`` (``TranslateSingleSeq``), kept because it is part of the
reference's importable surface and pinned by the CPU tests.  It is NOT on the product path: frames are translated on the
GPU, both for scoring (the kernels) and for the peptides the CLI emits (``orph_translate_frames`` / ``orph_stream_emit``).

Codon semantics (constants_translate.py:59-141): the 64 upper-case ACGT codons, the seven ``xxN``
codons ACN CTN CGN GTN GCN GGN TCN, everything else -> ``X``; reverse complement maps only ACGT.
"""
_BASES = "TCAG"
_AMINO = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"

CODON_TABLE = {a + b + c: _AMINO[16 * i + 4 * j + k]
               for i, a in enumerate(_BASES) for j, b in enumerate(_BASES) for k, c in enumerate(_BASES)}
CODON_TABLE.update({"ACN": "T", "CTN": "L", "CGN": "R", "GTN": "V", "GCN": "A", "GGN": "G", "TCN": "S"})

_COMPLEMENT = str.maketrans("ACGT", "TGCA")


class TranslateSingleSeq:
    def __init__(self, seq, verbose=False):
        self.seq = seq
        self.verbose = verbose
        self.sign = 1

    @staticmethod
    def _translate_codon(codon):
        return CODON_TABLE.get(codon, "X") if len(codon) == 3 else ""

    def _single_seq_translation(self, seq, frame=0):
        get = CODON_TABLE.get
        n = (len(seq) - frame) // 3 if len(seq) >= frame else 0
        return "".join(get(seq[frame + 3 * i: frame + 3 * i + 3], "X") for i in range(n))

    @staticmethod
    def _reverse_complement(seq):
        return seq.translate(_COMPLEMENT)[::-1]

    def three_frame_translation(self):
        strand = self.seq if self.sign == 1 else self._reverse_complement(self.seq)
        for frame in range(3):
            yield self._single_seq_translation(strand, frame)

    def three_frame_translation_stops(self, sign):
        self.sign = sign
        return {sign * (i + 1): t for i, t in enumerate(self.three_frame_translation())}

    def three_frame_translation_no_stops(self, sign):
        return {f: t for f, t in self.three_frame_translation_stops(sign).items() if "*" not in t}

    def six_frame_translation(self):
        out = self.three_frame_translation_stops(1)
        out.update(self.three_frame_translation_stops(-1))
        return out

    def six_frame_translation_no_stops(self):
        out = self.three_frame_translation_no_stops(1)
        out.update(self.three_frame_translation_no_stops(-1))
        return out

    def frame_translation(self, frame):
        """One frame by its key (1,2,3,-1,-2,-3)."""
        strand = self.seq if frame > 0 else self._reverse_complement(self.seq)
        return self._single_seq_translation(strand, abs(frame) - 1)
