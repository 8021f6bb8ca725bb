"""Reduced amino-acid alphabets as 256-entry byte tables.

Host-side mirror of the reference's ``orpheum/sequence_encodings.py`` for the names the
translate/index path uses: ``encode_peptide`` (``:457-468``), ``PEPTIDE_ENCODINGS``
(``:307-320``), ``VALID_PEPTIDE_MOLECULES`` (``:331-348``), ``ALPHABET_ALIASES`` (``:362-372``),
``ALPHABET_SIZES`` (``:384-400``), ``BEST_KSIZES`` (``:402-419``).  On the GPU an alphabet is only
ever a LUT[256] (``encode_lut``): unmapped bytes pass through unchanged, exactly like
``str.translate`` with a partial table.
"""
import numpy as np

# alphabet -> "letter:members" groups (dayhoff :32-58, dayhoff_v2 :62-88, hp :95-118, gbmr4 :133-155,
# sdm12 :164-185, hsdm17 :194-215, aa9 :221-242, botvinnik :245-274)
_GROUP_SPECS = {
    "dayhoff": "a:C b:AGPST c:DENQ d:HKR e:ILMV f:FWY",
    "dayhoff_v2": "a:C b:AGP B:ST c:DENQ d:HKR e:ILMV f:FWY",
    "hp": "h:AFGILMPVWY p:CDEHKNQRST",
    "gbmr4": "a:ADKERNTSQ b:YFLIVMCWH c:G d:P",
    "sdm12": "a:A b:D c:KER d:N e:TSQ f:YF g:LIVM h:C i:W j:H k:G l:P",
    "hsdm17": "a:A b:D c:KE d:R e:N f:T g:S h:Q i:Y j:F k:LIV l:M m:C n:G o:P p:W q:H",
    "aa9": "a:AST b:CFILMVY c:DN d:EQ e:G f:H g:KR h:P i:W",
    "botvinnik": "a:AG b:LIV c:FY d:ST e:NQ f:DE g:RK h:C i:M j:W k:H m:P",
}


def _mapping(spec):
    out = {}
    for group in spec.split():
        letter, members = group.split(":")
        for aa in members:
            out[aa] = letter
    return out


HP_MAPPING = _mapping(_GROUP_SPECS["hp"])
DAYHOFF_MAPPING = _mapping(_GROUP_SPECS["dayhoff"])
DAYHOFF_V2_MAPPING = _mapping(_GROUP_SPECS["dayhoff_v2"])
BOTVINNIK_MAPPING = _mapping(_GROUP_SPECS["botvinnik"])
AA9_MAPPING = _mapping(_GROUP_SPECS["aa9"])
GBMR4_MAPPING = _mapping(_GROUP_SPECS["gbmr4"])
SDM12_MAPPING = _mapping(_GROUP_SPECS["sdm12"])
HSDM17_MAPPING = _mapping(_GROUP_SPECS["hsdm17"])

# alphabet name -> str.translate table
PEPTIDE_ENCODINGS = {
    name: str.maketrans(mapping)
    for name, mapping in (
        ("hp", HP_MAPPING), ("hp2", HP_MAPPING), ("hydrophobic-polar", HP_MAPPING),
        ("dayhoff", DAYHOFF_MAPPING), ("dayhoff6", DAYHOFF_MAPPING), ("dayhoff_v2", DAYHOFF_V2_MAPPING),
        ("botvinnik", BOTVINNIK_MAPPING), ("botvinnik8", BOTVINNIK_MAPPING),
        ("aa9", AA9_MAPPING), ("gbmr4", GBMR4_MAPPING), ("sdm12", SDM12_MAPPING), ("hsdm17", HSDM17_MAPPING),
    )
}

PROTEIN_LIKE = ("protein", "peptide", "protein20", "peptide20", "aa20")
VALID_PEPTIDE_MOLECULES = PROTEIN_LIKE + (
    "dayhoff", "dayhoff6", "botvinnik", "botvinnik8", "hydrophobic-polar", "hp", "hp2", "aa9", "gbmr4", "sdm12",
    "hsdm17",
)

ALPHABET_ALIASES = {
    "protein20": ("peptide", "protein", "aa20", "protein20"),
    "dayhoff6": ("dayhoff", "dayhoff6"),
    "hp2": ("hydrophobic-polar", "hp", "hp2"),
    "botvinnik8": ("botvinnik", "botvinnik8"),
    "aa9": ("aa9",),
    "gbmr4": ("gbmr4",),
    "sdm12": ("sdm12",),
    "hsdm17": ("hsdm17",),
}
ALIAS_TO_ALPHABET = {alias: canonical for canonical, aliases in ALPHABET_ALIASES.items() for alias in aliases}

_SIZES = {"protein20": 20, "dayhoff6": 6, "botvinnik8": 8, "hp2": 2, "aa9": 9, "gbmr4": 4, "sdm12": 12, "hsdm17": 17}
_KSIZES = {"protein20": 7, "dayhoff6": 12, "botvinnik8": 11, "hp2": 31, "aa9": 10, "gbmr4": 16, "sdm12": 9, "hsdm17": 8}
ALPHABET_SIZES = {alias: _SIZES[canon] for alias, canon in ALIAS_TO_ALPHABET.items()}
ALPHABET_SIZES["peptide20"] = 20
BEST_KSIZES = {alias: _KSIZES[canon] for alias, canon in ALIAS_TO_ALPHABET.items()}
BEST_KSIZES["peptide20"] = 7


def encode_peptide(peptide_sequence, molecule):
    """Re-encode a peptide string; identity for the protein-like names; ValueError otherwise."""
    table = PEPTIDE_ENCODINGS.get(molecule)
    if table is not None:
        return peptide_sequence.translate(table)
    if molecule in VALID_PEPTIDE_MOLECULES:
        return peptide_sequence
    raise ValueError(
        f"{molecule} is not a valid amino acid encoding, only {', '.join(PEPTIDE_ENCODINGS.keys())} can be used"
    )


def encode_lut(molecule):
    """The same mapping as a uint8[256] table for the kernels."""
    lut = np.arange(256, dtype=np.uint8)
    table = PEPTIDE_ENCODINGS.get(molecule)
    if table is not None:
        for src, dst in table.items():
            lut[src] = ord(dst) if isinstance(dst, str) else dst
    elif molecule not in VALID_PEPTIDE_MOLECULES:
        raise ValueError(
            f"{molecule} is not a valid amino acid encoding, only {', '.join(PEPTIDE_ENCODINGS.keys())} can be used"
        )
    return lut
