/* tests/pg_stub: empty stand-in for catalog/pg_type.h (see postgres.h) */
