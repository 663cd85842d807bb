/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/field.h header).
 *
 * SHA-256 (FIPS 180-4), the algorithm behind the reference's `sha2` 0.10.8
 * dependency (Cargo.lock:261-269; call sites src/merkle_tree/mod.rs:112-114,
 * 178-189 and src/transcript.rs:19-32).  Streaming init/update/final mirrors
 * `Sha256::new(); update(); finalize()`.  Like the crate (via `cpufeatures`)
 * it uses the x86 SHA extensions when the CPU has them, else portable C.
 */
#ifndef ORACLE_SHA256_H
#define ORACLE_SHA256_H
#include <stddef.h>
#include <stdint.h>

typedef struct {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len; /* total bytes absorbed */
} or_sha256_ctx;

void or_sha256_init(or_sha256_ctx *c);
void or_sha256_update(or_sha256_ctx *c, const void *data, size_t len);
void or_sha256_final(or_sha256_ctx *c, uint8_t out[32]); /* consumes c */
void or_sha256(const void *data, size_t len, uint8_t out[32]);
int or_sha256_uses_shani(void);
void or_sha256_force_portable(int on);

#endif
