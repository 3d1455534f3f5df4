"""CPU oracle of the Overcooked hot path — TEST INFRASTRUCTURE ONLY (see ovc_oracle.c header)."""
