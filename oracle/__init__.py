"""CPU oracle of the PTGEnv hot path -- TEST INFRASTRUCTURE ONLY (never imported by rl_ptg_b200)."""
