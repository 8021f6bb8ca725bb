"""Stand-in for sourmash (pinned sourmash==3.2.2 in the reference's requirements_with_versions.txt:22)."""
