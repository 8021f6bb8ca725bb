from .commandline import cli

cli()
