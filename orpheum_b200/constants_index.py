"""Bloom-index defaults and the ``--tablesize`` option type.

Values mirror the reference (``orpheum/constants_index.py:4-10``): four tables of 1e8 bits, k = 7 / 12 / 31 for
the protein / dayhoff / hydrophobic-polar alphabets.  ``BASED_INT`` lets the command line say ``1e8`` as well
as ``100000000`` (reference behaviour: ``int(mantissa * 10 ** exponent)``, ``:14-35``).
"""
import click

DEFAULT_N_TABLES, DEFAULT_MAX_TABLESIZE = 4, 10 ** 8
DEFAULT_PROTEIN_KSIZE, DEFAULT_DAYHOFF_KSIZE, DEFAULT_HP_KSIZE = 7, 12, 31


def parse_based_int(text):
    """'12345' -> 12345, '1e8' -> 100000000, '2.5e3' -> 2500; ValueError / TypeError otherwise."""
    text = str(text)
    mantissa, sep, exponent = text.partition("e")
    return int(float(mantissa) * 10 ** int(exponent)) if sep else int(text, 10)


class BasedIntParamType(click.ParamType):
    name = "integer"

    def convert(self, value, param, ctx):
        if isinstance(value, int):
            return value
        try:
            return parse_based_int(value)
        except (TypeError, ValueError):
            self.fail(f"{value!r} is not a valid integer", param, ctx)


BASED_INT = BasedIntParamType()
