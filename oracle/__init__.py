"""CPU oracle for the orpheum translate/index hot path -- test infrastructure only.

Nothing under ``orpheum_b200/`` may import this package.  See orpheum_oracle.c.
"""
