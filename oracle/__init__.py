"""CPU oracle package — TEST INFRASTRUCTURE ONLY (see oracle_core.hpp). Never imported by posidonius_b200/."""
