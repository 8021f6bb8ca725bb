"""Defaults of the Bloom index (mirror of the reference's ``orpheum/constants_index.py:4-38``)."""
import click

DEFAULT_N_TABLES = 4
DEFAULT_MAX_TABLESIZE = int(1e8)

DEFAULT_PROTEIN_KSIZE = 7
DEFAULT_DAYHOFF_KSIZE = 12
DEFAULT_HP_KSIZE = 31


class BasedIntParamType(click.ParamType):
    """click type that accepts ``100000000`` as well as ``1e8`` (reference: constants_index.py:14-35)."""

    name = "integer"

    def convert(self, value, param, ctx):
        if isinstance(value, int):
            return value
        try:
            text = str(value)
            if "e" in text:
                mantissa, exponent = text.split("e")
                return int(float(mantissa) * 10 ** int(exponent))
            return int(text, 10)
        except (TypeError, ValueError):
            self.fail(f"{value!r} is not a valid integer", param, ctx)


BASED_INT = BasedIntParamType()
