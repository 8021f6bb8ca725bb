"""``orpheum`` command group: ``translate`` and ``index`` (mirror of ``orpheum/commandline.py:21-36``;
``compare-kmers`` is a separate analysis tool outside the scoring path and is not provided)."""
from functools import partial

import click

click.option = partial(click.option, show_default=True)  # as the reference does (commandline.py:16)

from .index import cli as index  # noqa: E402
from .translate import cli as translate  # noqa: E402

settings = dict(help_option_names=["-h", "--help"])


@click.group(options_metavar="", subcommand_metavar="<command>", context_settings=settings)
def cli():
    """
    B200-native orpheum: partition sequencing reads into protein-coding and non-coding bins
    (six-frame translation + k-mer Bloom filter of a reference proteome), scored on the GPU.
    """


cli.add_command(index, name="index")
cli.add_command(translate, name="translate")

if __name__ == "__main__":
    cli()
