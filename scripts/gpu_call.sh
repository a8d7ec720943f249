timeout 300 python scripts/prof_commit.py --start random
timeout 300 python scripts/prof_commit.py --start truth10
