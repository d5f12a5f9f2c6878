// The remaining eval.py metrics on the device, built on the kNN / Chamfer kernels of this library:
//
//   normals_pca_kernel   Open3D's estimate_normals(KDTreeSearchParamKNN(knn)) as eval.py:58-59 calls it: per point, the
//                        covariance of its knn nearest points (itself included) and the eigenvector of the smallest
//                        eigenvalue.  Double precision; closed-form eigenvalues of the symmetric 3x3 matrix, eigenvector by
//                        the best-conditioned cross product of two rows of (C - lambda I).  The sign of a normal is
//                        arbitrary and irrelevant to eval.py, which only squares the projection.
//   p2plane_kernel       eval.py:68-92: for every reconstructed point, its nearest original point (index from the Chamfer /
//                        1-NN kernel), e = (diff . normal)^2, mse = mean e, psnr = 10 log10(|bbox diag|^2 / mse).
//   uc_kernel            eval.py:127-151 calc_uc, last two lines: variance (population, np.var) of the self-neighbour
//                        distances of the decompressed region over that of the input region.
// All reductions are fixed-order double sums, one CTA per cloud.
#include "pcc_common.cuh"

namespace pcc {

// grid: ceil(rows / 128); nn [rows, knn, 3] = the knn nearest points of each point (pcc_knn_f32's out_nn, no recentring)
__global__ void __launch_bounds__(128)
normals_pca_kernel(const float *__restrict__ nn, long long rows, int knn, float *__restrict__ normals) {
    const long long r = blockIdx.x * 128ll + threadIdx.x;
    if (r >= rows) return;
    const float *p = nn + r * knn * 3;
    double sx = 0, sy = 0, sz = 0, xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
    for (int k = 0; k < knn; ++k) {
        const double x = p[k * 3 + 0], y = p[k * 3 + 1], z = p[k * 3 + 2];
        sx += x, sy += y, sz += z;
        xx += x * x, xy += x * y, xz += x * z, yy += y * y, yz += y * z, zz += z * z;
    }
    const double inv = 1.0 / knn;
    const double mx = sx * inv, my = sy * inv, mz = sz * inv;
    double a00 = xx * inv - mx * mx, a01 = xy * inv - mx * my, a02 = xz * inv - mx * mz;
    double a11 = yy * inv - my * my, a12 = yz * inv - my * mz, a22 = zz * inv - mz * mz;
    // scale to O(1) so the closed form is well conditioned
    const double s = fmax(fmax(fmax(fabs(a00), fabs(a11)), fmax(fabs(a22), fabs(a01))), fmax(fabs(a02), fabs(a12)));
    double nx = 0.0, ny = 0.0, nz = 1.0;   // Open3D's fallback for a degenerate neighbourhood
    if (s > 0.0) {
        const double is = 1.0 / s;
        a00 *= is, a01 *= is, a02 *= is, a11 *= is, a12 *= is, a22 *= is;
        // smallest eigenvalue of the symmetric matrix (trigonometric solution of the characteristic cubic)
        const double q = (a00 + a11 + a22) / 3.0;
        const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
        const double p2 = (b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * (a01 * a01 + a02 * a02 + a12 * a12)) / 6.0;
        const double pp = sqrt(p2);
        double lam = q;
        if (pp > 1e-300) {
            const double ip = 1.0 / pp;
            const double c00 = b00 * ip, c01 = a01 * ip, c02 = a02 * ip, c11 = b11 * ip, c12 = a12 * ip, c22 = b22 * ip;
            double half_det = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
            half_det = fmin(1.0, fmax(-1.0, half_det));
            const double phi = acos(half_det) / 3.0;
            lam = q + 2.0 * pp * cos(phi + 2.0943951023931953);   // + 2 pi / 3: the smallest of the three roots
        }
        // eigenvector: rows of (A - lam I) are orthogonal to it; take the largest of the three pairwise cross products
        const double r0x = a00 - lam, r0y = a01, r0z = a02, r1x = a01, r1y = a11 - lam, r1z = a12, r2x = a02, r2y = a12, r2z = a22 - lam;
        const double c0x = r0y * r1z - r0z * r1y, c0y = r0z * r1x - r0x * r1z, c0z = r0x * r1y - r0y * r1x;
        const double c1x = r0y * r2z - r0z * r2y, c1y = r0z * r2x - r0x * r2z, c1z = r0x * r2y - r0y * r2x;
        const double c2x = r1y * r2z - r1z * r2y, c2y = r1z * r2x - r1x * r2z, c2z = r1x * r2y - r1y * r2x;
        const double n0 = c0x * c0x + c0y * c0y + c0z * c0z, n1 = c1x * c1x + c1y * c1y + c1z * c1z, n2 = c2x * c2x + c2y * c2y + c2z * c2z;
        double vx = c0x, vy = c0y, vz = c0z, nn2 = n0;
        if (n1 > nn2) vx = c1x, vy = c1y, vz = c1z, nn2 = n1;
        if (n2 > nn2) vx = c2x, vy = c2y, vz = c2z, nn2 = n2;
        if (nn2 > 0.0) {
            const double in = rsqrt(nn2);
            nx = vx * in, ny = vy * in, nz = vz * in;
        }
    }
    normals[r * 3 + 0] = static_cast<float>(nx);
    normals[r * 3 + 1] = static_cast<float>(ny);
    normals[r * 3 + 2] = static_cast<float>(nz);
}

__device__ __forceinline__ double block_sum_256(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    __syncthreads();
    return t;
}

// grid B; out [B, 2] = (p2plane mse, p2plane psnr dB)
__global__ void __launch_bounds__(256)
p2plane_kernel(const float *__restrict__ recon, const float *__restrict__ orig, const int64_t *__restrict__ ix,
               const float *__restrict__ normals, const float *__restrict__ bbox, int P1, int P2, double *__restrict__ out) {
    __shared__ double red[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *rc = recon + static_cast<size_t>(b) * P1 * 3, *og = orig + static_cast<size_t>(b) * P2 * 3;
    const float *nm = normals + static_cast<size_t>(b) * P2 * 3;
    double s = 0.0;
    for (int i = tid; i < P1; i += 256) {
        const long long j = ix[static_cast<size_t>(b) * P1 + i];
        const double dx = static_cast<double>(rc[i * 3 + 0]) - static_cast<double>(og[j * 3 + 0]);
        const double dy = static_cast<double>(rc[i * 3 + 1]) - static_cast<double>(og[j * 3 + 1]);
        const double dz = static_cast<double>(rc[i * 3 + 2]) - static_cast<double>(og[j * 3 + 2]);
        const double d = dx * nm[j * 3 + 0] + dy * nm[j * 3 + 1] + dz * nm[j * 3 + 2];
        s += d * d;
    }
    const double t = block_sum_256(s, red);
    if (tid == 0) {
        const double mse = t / static_cast<double>(P1);
        const float *bb = bbox + static_cast<size_t>(b) * 6;
        double diag2 = 0.0;
        for (int a = 0; a < 3; ++a) {
            const double e = static_cast<double>(bb[3 + a]) - static_cast<double>(bb[a]);
            diag2 += e * e;
        }
        out[b * 2 + 0] = mse;
        out[b * 2 + 1] = mse > 0.0 ? 10.0 * log10(diag2 / mse) : INFINITY;
    }
}

// grid B; d2_in / d2_dec [B, n] = squared distance of every region point to its nearest other region point
__global__ void __launch_bounds__(256)
uc_kernel(const float *__restrict__ d2_in, const float *__restrict__ d2_dec, int n, double *__restrict__ out) {
    __shared__ double red[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    double var[2];
    for (int side = 0; side < 2; ++side) {
        const float *d = (side ? d2_dec : d2_in) + static_cast<size_t>(b) * n;
        double s = 0.0;
        for (int i = tid; i < n; i += 256) s += sqrt(static_cast<double>(d[i]));
        const double mean = block_sum_256(s, red) / n;
        double v = 0.0;
        for (int i = tid; i < n; i += 256) {
            const double e = sqrt(static_cast<double>(d[i])) - mean;
            v += e * e;
        }
        var[side] = block_sum_256(v, red) / n;
    }
    if (tid == 0) out[b] = var[1] / var[0];
}

}  // namespace pcc

PCC_API int pcc_normals_pca_f32(const float *nn, int64_t rows, int knn, float *out_normals, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(nn && out_normals, "pcc_normals_pca_f32: null pointer");
    PCC_REQUIRE(rows >= 0 && knn >= 3, "pcc_normals_pca_f32: need knn >= 3 (got %d)", knn);
    if (rows == 0) return 0;
    normals_pca_kernel<<<static_cast<unsigned>((rows + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(nn, rows, knn, out_normals);
    return check_launch("normals_pca_kernel");
}

PCC_API int pcc_p2plane_f32(const float *recon, const float *orig, const int64_t *ix, const float *normals, const float *bbox, int B,
                            int P1, int P2, double *out, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(recon && orig && ix && normals && bbox && out, "pcc_p2plane_f32: null pointer");
    PCC_REQUIRE(B >= 1 && P1 >= 1 && P2 >= 1, "pcc_p2plane_f32: bad shape");
    p2plane_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(recon, orig, ix, normals, bbox, P1, P2, out);
    return check_launch("p2plane_kernel");
}

PCC_API int pcc_uc_f32(const float *d2_in, const float *d2_dec, int B, int n, double *out, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(d2_in && d2_dec && out, "pcc_uc_f32: null pointer");
    PCC_REQUIRE(B >= 1 && n >= 2, "pcc_uc_f32: bad shape");
    uc_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(d2_in, d2_dec, n, out);
    return check_launch("uc_kernel");
}
