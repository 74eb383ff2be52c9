"""ORACLE package: CPU restatements used ONLY as the checker (tests/, smoke(), bench.py's
cpu_baseline / --impl reference legs).  Nothing under uav-airvision_b200/ imports it."""
