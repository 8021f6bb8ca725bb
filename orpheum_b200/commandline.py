"""``orpheum`` command group with the two sub-commands on the GPU path, ``translate`` and ``index``
(same names, options and help behaviour as ``orpheum/commandline.py:21-36``; ``compare-kmers`` is a separate
analysis tool outside the scoring path and is not provided)."""
import click

from . import index as _index
from . import translate as _translate

# every option shows its default in --help, as in the reference's command line (commandline.py:16 patches click.option for
# the whole process; here only the options of these two commands are touched, whatever the import order)
for _command in (_index.cli, _translate.cli):
    for _param in _command.params:
        if isinstance(_param, click.Option) and _param.show_default is None:
            _param.show_default = True

cli = click.Group(
    name="orpheum",
    help="B200-native orpheum: partition sequencing reads into protein-coding and non-coding bins "
         "(six-frame translation + k-mer Bloom filter of a reference proteome), scored on the GPU.",
    commands={"index": _index.cli, "translate": _translate.cli},
    options_metavar="",
    subcommand_metavar="<command>",
    context_settings={"help_option_names": ["-h", "--help"]},
)

if __name__ == "__main__":
    cli()
