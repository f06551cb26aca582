bash profiles/scripts/r2_run4.sh
bash profiles/scripts/r2_run5.sh
