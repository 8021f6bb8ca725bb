"""orpheum_b200: the scoring hot path of czbiohub-sf/orpheum (`translate` + `index`) on NVIDIA B200.

Nothing heavy is imported eagerly (the reference's ``orpheum/__init__.py`` pulls in notebook and
plotting helpers; this package deliberately does not).  Sub-modules:

    translate, index, commandline   the reference's host interface for this path (click CLI, Translate)
    gpu, _lib                        C-ABI binding: Context, GpuNodegraph, score_reads
    sequence_encodings, constants_*  alphabets and output strings
    create_save_summary, fastx       output writers, FASTA/FASTQ reader
    build, synth                     nvcc build of csrc/, synthetic workloads for tests and bench.py
"""
__version__ = "0.1.0"
