// mesh.cuh -- host-side mesh preparation (validation, sign fields, locality renumbering, SoA
// transposition) and the device-resident mesh.
//
// Stands behind ReadHorzMesh / signIndexField! / Adapt.adapt_structure(backend, mesh)
// (reference src/infra/MPASMesh/HorzMesh.jl:292-355) and VerticalMesh (VertMesh.jl:46-82).
//
// Device layout (all 0-based, renumbered):
//   * connectivity is Int32 and slot-major ("SoA-transposed"): row i of edgesOnEdge is a contiguous
//     array over edges, so a warp reading slot i of 32 consecutive edges reads 128 contiguous bytes.
//   * cells follow a Hilbert curve over (xCell, yCell) (Morton over x,y,z when z varies); edges are
//     ordered by their owner cell cellsOnEdge[1] (stable), vertices by cellsOnVertex-free rule:
//     the smallest new edge id touching them.  Index ranges are therefore compact in space and the
//     indirect gathers of neighbouring threads fall into the same few cache lines.
#pragma once
#include <algorithm>
#include <cmath>
#include <numeric>

#include "common.cuh"

namespace mokab {

// Hilbert index of (x, y) on a 2^order grid.
static inline uint64_t hilbert_xy(uint32_t x, uint32_t y, int order)
{
    uint64_t d = 0;
    for (uint32_t s = 1u << (order - 1); s > 0; s >>= 1) {
        uint32_t rx = (x & s) ? 1 : 0, ry = (y & s) ? 1 : 0;
        d += (uint64_t)s * s * ((3 * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) {
                x = s - 1 - x;
                y = s - 1 - y;
            }
            std::swap(x, y);
        }
    }
    return d;
}

static inline uint64_t spread3(uint64_t v)  // 21 bits -> every third bit
{
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}

// Everything the kernels need, on the host, renumbered and 0-based (-1 = absent).
struct HostMesh {
    int64_t nC = 0, nE = 0, nV = 0;
    int64_t nCo = 0, nEo = 0;                  // owned (computed) cells / edges; the rest are halo copies
    std::vector<int32_t> blkEdgeStart, blkInterior, blkBoundary;  // fused-kernel blocks (kBlockCells cells each)
    // derived edgesOnEdge (see below): position of each edge in the edgesOnCell rows of its two cells, and the
    // blocks in which every edge's edgesOnEdge row equals the row derived from edgesOnCell
    std::vector<uint8_t> posE, blkDerived;
    int S = 0, S2 = 0, D = 0;  // maxEdges, maxEdges2, vertexDegree
    std::vector<int32_t> permC, permE, permV;  // perm[new] = old
    std::vector<int32_t> ce;                   // (nE, 2) c1, c2 (c2 = c1 on masked edges)
    std::vector<int32_t> eoe;                  // slot-major (S2, nE)
    std::vector<double> woe;                   // slot-major (S2, nE); 0 beyond nEoE / on masked edges
    std::vector<uint8_t> nEoE;
    // decomposed meshes: the edgesOnEdge / weightsOnEdge rows of the HALO edges, edge-major (nE - nEo, haloS2), new numbering, -1 =
    // absent or not local.  No forward kernel reads them (a halo edge's tendency is its owner's business); the reverse mode does:
    // the transposed Coriolis stencil of an owned edge lists every edge whose sum reads it, halo edges included.
    std::vector<int32_t> haloEoe;
    std::vector<double> haloWoe;
    int haloS2 = 0;
    std::vector<double> dc, dv, fE;
    std::vector<int32_t> eoc, sgnC;            // slot-major (S, nC)
    std::vector<uint8_t> nEoC;
    std::vector<double> area, H;
    std::vector<int32_t> eov, sgnV;            // slot-major (D, nV)
    std::vector<double> areaTri;
    bool any_boundary = false;
    bool uniformF = true;     // fEdge identical on every edge (f-plane): Coriolis weights stay unfolded
    double f0 = 0.0;
};

constexpr int kBlockCells = MOKAB_BLOCK_CELLS;  // cells per block of the fused kernel (fused::kTC)

static void build_host_mesh(const mokab_mesh_desc &d, uint32_t flags, HostMesh &m)
{
    MOKAB_REQUIRE(d.nCells > 0 && d.nEdges > 0, "mesh_create: nCells and nEdges must be positive");
    MOKAB_REQUIRE(d.nEdges < (int64_t)1 << 30 && d.nCells < (int64_t)1 << 30, "mesh_create: mesh too large for Int32 ids");
    MOKAB_REQUIRE(d.maxEdges > 0 && d.maxEdges <= 16 && d.maxEdges2 > 0 && d.maxEdges2 <= 32,
                  "mesh_create: maxEdges must be in 1..16 and maxEdges2 in 1..32");
    MOKAB_REQUIRE(d.cellsOnEdge && d.edgesOnEdge && d.nEdgesOnEdge && d.weightsOnEdge && d.dcEdge && d.dvEdge,
                  "mesh_create: missing Edges arrays");
    MOKAB_REQUIRE(d.edgesOnCell && d.nEdgesOnCell && d.areaCell, "mesh_create: missing PrimaryCells arrays");
    MOKAB_REQUIRE(d.restingThicknessSum, "mesh_create: missing restingThicknessSum");
    if (d.nVertices > 0)
        MOKAB_REQUIRE(d.edgesOnVertex && d.areaTriangle && d.vertexDegree > 0 && d.vertexDegree <= 8 &&
                          (d.edgeSignOnVertex || d.verticesOnEdge),
                      "mesh_create: nVertices > 0 needs edgesOnVertex, areaTriangle and edgeSignOnVertex or verticesOnEdge");

    const int64_t nC = d.nCells, nE = d.nEdges, nV = d.nVertices;
    const int Sf = (int)d.maxEdges, S2f = (int)d.maxEdges2, D = nV ? (int)d.vertexDegree : 0;   // row widths of the caller's arrays
    const int64_t nCo = d.nCellsOwned > 0 ? d.nCellsOwned : nC, nEo = d.nEdgesOwned > 0 ? d.nEdgesOwned : nE;
    MOKAB_REQUIRE(nCo <= nC && nEo <= nE, "mesh_create: nCellsOwned/nEdgesOwned exceed nCells/nEdges");
    // Device rows are as wide as the longest live row, not as the file's maxEdges / maxEdges2 (MPAS files often
    // carry padding): a mesh of hexagons stored with maxEdges = 7 still gets the compile-time (10, 6) kernels.
    int S = Sf, S2 = S2f;
    if (!(flags & MOKAB_MESH_KEEP_WIDTHS)) {
        int s = 1, s2 = 1;
        for (int64_t c = 0; c < nCo; ++c) s = std::max(s, std::min((int)d.nEdgesOnCell[c], Sf));
        for (int64_t e = 0; e < nEo; ++e) s2 = std::max(s2, std::min((int)d.nEdgesOnEdge[e], S2f));
        // ... rounded up to the next pair the compile-time kernels exist for (padding slots: index = self, weight 0, sign 0), so
        // that e.g. a part of a decomposed pentagon / hexagon / heptagon mesh whose longest rows happen to be (11, 7) gets the
        // (12, 7) kernels like its neighbours, and a mesh of squares the (10, 6) ones
        if (s <= 6 && s2 <= 10) { s = 6; s2 = 10; }
        else if (s <= 7 && s2 <= 12) { s = 7; s2 = 12; }
        S = s; S2 = s2;
    }
    MOKAB_REQUIRE((nCo == nC && nEo == nE) || nV == 0, "mesh_create: decomposed meshes carry no vertex arrays");
    m.nC = nC; m.nE = nE; m.nV = nV; m.S = S; m.S2 = S2; m.D = D; m.nCo = nCo; m.nEo = nEo;

    // ---- validate index ranges (the reference does none; a bad index here would fault the GPU) ----
    for (int64_t e = nEo; e < nE; ++e) {  // halo edges: at least one local cell, first cell not owned
        int32_t c1 = d.cellsOnEdge[2 * e], c2 = d.cellsOnEdge[2 * e + 1];
        MOKAB_REQUIRE(c1 >= 0 && c1 <= nC && c2 >= 0 && c2 <= nC && (c1 > 0 || c2 > 0), "mesh_create: halo edge without a local cell");
        MOKAB_REQUIRE(c1 == 0 || c1 > nCo, "mesh_create: an edge whose first cell is owned must be among the first nEdgesOwned edges");
    }
    for (int64_t e = 0; e < nEo; ++e) {
        int32_t c1 = d.cellsOnEdge[2 * e], c2 = d.cellsOnEdge[2 * e + 1];
        bool bnd = d.boundaryEdge && d.boundaryEdge[e];
        MOKAB_REQUIRE(c1 >= 1 && c1 <= nCo, "mesh_create: cellsOnEdge[1, e] out of range (must be an owned cell)");
        MOKAB_REQUIRE((c2 >= 1 && c2 <= nC) || (bnd && c2 == 0),
                      "mesh_create: cellsOnEdge[2, e] out of range (0 is allowed only on boundaryEdge edges)");
        int32_t n = d.nEdgesOnEdge[e];
        MOKAB_REQUIRE(n >= 0 && n <= S2f, "mesh_create: nEdgesOnEdge out of range");
        for (int i = 0; i < n; ++i) {
            int32_t x = d.edgesOnEdge[(int64_t)S2f * e + i];
            MOKAB_REQUIRE(x >= 0 && x <= nE, "mesh_create: edgesOnEdge out of range");
        }
        if (bnd) m.any_boundary = true;
    }
    for (int64_t c = 0; c < nCo; ++c) {
        int32_t n = d.nEdgesOnCell[c];
        MOKAB_REQUIRE(n >= 1 && n <= Sf, "mesh_create: nEdgesOnCell out of range");
        for (int i = 0; i < n; ++i) {
            int32_t x = d.edgesOnCell[(int64_t)Sf * c + i];
            MOKAB_REQUIRE(x >= 1 && x <= nE, "mesh_create: edgesOnCell out of range");
        }
    }
    for (int64_t v = 0; v < nV; ++v)
        for (int j = 0; j < D; ++j) {
            int32_t x = d.edgesOnVertex[(int64_t)D * v + j];
            MOKAB_REQUIRE(x >= 1 && x <= nE, "mesh_create: edgesOnVertex out of range");
        }

    // ---- cell permutation: space-filling curve over the cell centres ---------------------------
    m.permC.resize(nC);
    std::iota(m.permC.begin(), m.permC.end(), 0);
    if ((flags & MOKAB_MESH_RENUMBER) && d.xCell && d.yCell) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        const double *xyz[3] = {d.xCell, d.yCell, d.zCell};
        for (int a = 0; a < 3; ++a) {
            if (!xyz[a]) { lo[a] = hi[a] = 0; continue; }
            for (int64_t c = 0; c < nC; ++c) { lo[a] = std::min(lo[a], xyz[a][c]); hi[a] = std::max(hi[a], xyz[a][c]); }
        }
        const bool planar = !(hi[2] > lo[2]);
        std::vector<uint64_t> key(nC);
        if (planar) {
            const int order = 16;
            const double ext = std::max(hi[0] - lo[0], hi[1] - lo[1]);
            const double sc = ext > 0 ? (double)((1u << order) - 1) / ext : 0.0;
#pragma omp parallel for schedule(static)
            for (int64_t c = 0; c < nC; ++c) {
                uint32_t qx = (uint32_t)((d.xCell[c] - lo[0]) * sc), qy = (uint32_t)((d.yCell[c] - lo[1]) * sc);
                key[c] = hilbert_xy(qx, qy, order);
            }
        } else {
            double sc[3];
            for (int a = 0; a < 3; ++a) sc[a] = hi[a] > lo[a] ? 2097151.0 / (hi[a] - lo[a]) : 0.0;
#pragma omp parallel for schedule(static)
            for (int64_t c = 0; c < nC; ++c)
                key[c] = spread3((uint64_t)((d.xCell[c] - lo[0]) * sc[0])) | spread3((uint64_t)((d.yCell[c] - lo[1]) * sc[1])) << 1 |
                         spread3((uint64_t)((d.zCell[c] - lo[2]) * sc[2])) << 2;
        }
        auto by_key = [&](int32_t a, int32_t b) { return key[a] < key[b]; };
        std::stable_sort(m.permC.begin(), m.permC.begin() + nCo, by_key);  // owned first ...
        std::stable_sort(m.permC.begin() + nCo, m.permC.end(), by_key);    // ... halo copies after
    }
    std::vector<int32_t> invC(nC);
    for (int64_t i = 0; i < nC; ++i) invC[m.permC[i]] = (int32_t)i;

    // ---- edge permutation: counting sort by new id of the owner cell cellsOnEdge[1] (stable) --------
    m.permE.resize(nE);
    std::vector<int32_t> invE(nE);
    {
        // bucket = new id of the first cell; halo edges without a local first cell go last (bucket nC)
        auto bucket = [&](int64_t e) -> int64_t { int32_t c1 = d.cellsOnEdge[2 * e]; return c1 > 0 ? invC[c1 - 1] : nC; };
        std::vector<int64_t> start(nC + 2, 0);
        for (int64_t e = 0; e < nE; ++e) start[bucket(e) + 1]++;
        for (int64_t c = 0; c <= nC; ++c) start[c + 1] += start[c];
        for (int64_t e = 0; e < nE; ++e) {
            int64_t pos = start[bucket(e)]++;
            m.permE[pos] = (int32_t)e;
        }
        // ... then, inside every block of kBlockCells cells, SLOT-MAJOR: first the first edge every cell of the block owns
        // (in the order of its edgesOnCell row), then every cell's second edge, ...  Thread t of the fused kernels handles
        // edge blkEdgeStart + t + k * blockDim in its k-th iteration, so a warp now works on the SAME kind of edge of 32
        // consecutive cells (on a hexagon mesh: 32 east edges, then 32 north-east edges, ...): the neighbour edges and cells
        // it gathers are then consecutive too -- the ten normalVelocity gathers of the Coriolis sum, the edgesOnCell rows
        // the edgesOnEdge rebuild reads, and the per-slot gathers of the cell phase all become (nearly) coalesced, instead
        // of striding by the three edges a cell owns.  ncu (profiles/README.md, r02a): the stage kernel ran at 71 % of the L1
        // data-pipe wavefront limit against 66 % of DRAM, i.e. it was paying for scattered gathers, not for bytes.
        if (!(flags & MOKAB_MESH_EDGES_BY_CELL)) {
            std::vector<uint8_t> rank(nE, 0);
#pragma omp parallel for schedule(static)
            for (int64_t co = 0; co < nCo; ++co) {
                int r = 0;
                const int n = std::min<int>(d.nEdgesOnCell[co], Sf);
                for (int i = 0; i < n; ++i) {
                    const int64_t e = (int64_t)d.edgesOnCell[(int64_t)Sf * co + i] - 1;
                    if (e >= 0 && e < nE && d.cellsOnEdge[2 * e] == (int32_t)(co + 1)) rank[e] = (uint8_t)(r++);
                }
            }
            const int64_t nblocks = (nCo + kBlockCells - 1) / kBlockCells;
#pragma omp parallel for schedule(dynamic, 64)
            for (int64_t b = 0; b < nblocks; ++b) {
                // the block's edges are the contiguous run of the cell-sorted order whose owner lies in the block
                const int64_t c0 = b * kBlockCells, c1 = std::min<int64_t>(nCo, c0 + kBlockCells);
                const int64_t e0 = c0 == 0 ? 0 : start[c0 - 1], e1 = start[c1 - 1];
                std::stable_sort(m.permE.begin() + e0, m.permE.begin() + e1,
                                 [&](int32_t a, int32_t bb) { return rank[a] < rank[bb]; });
            }
        }
        for (int64_t i = 0; i < nE; ++i) invE[m.permE[i]] = (int32_t)i;
    }
    // ---- vertex permutation: by the smallest new edge id on the vertex ------------------------------
    std::vector<int32_t> invV(nV);
    m.permV.resize(nV);
    if (nV) {
        std::vector<int64_t> key(nV);
        for (int64_t v = 0; v < nV; ++v) {
            int64_t k = INT64_MAX;
            for (int j = 0; j < D; ++j) k = std::min<int64_t>(k, invE[d.edgesOnVertex[(int64_t)D * v + j] - 1]);
            key[v] = k;
        }
        std::iota(m.permV.begin(), m.permV.end(), 0);
        std::stable_sort(m.permV.begin(), m.permV.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
        for (int64_t i = 0; i < nV; ++i) invV[m.permV[i]] = (int32_t)i;
    }

    // ---- edges -----------------------------------------------------------------------------------
    m.ce.resize(2 * nE); m.eoe.assign((size_t)S2 * nE, -1); m.woe.assign((size_t)S2 * nE, 0.0);
    m.nEoE.resize(nE); m.dc.resize(nE); m.dv.resize(nE); m.fE.resize(nE);
    m.haloS2 = S2f;
    m.haloEoe.assign((size_t)(nE - nEo) * S2f, -1); m.haloWoe.assign((size_t)(nE - nEo) * S2f, 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t en = 0; en < nE; ++en) {
        const int64_t eo = m.permE[en];
        const bool bnd = d.boundaryEdge && d.boundaryEdge[eo];
        int32_t c1 = d.cellsOnEdge[2 * eo], c2 = d.cellsOnEdge[2 * eo + 1];
        if (c1 < 1) c1 = c2;                                                   // halo edge, first cell not local
        m.ce[2 * en] = invC[c1 - 1];
        m.ce[2 * en + 1] = (c2 >= 1 && !bnd) ? invC[c2 - 1] : invC[c1 - 1];  // masked: zero gradient
        m.dc[en] = d.dcEdge[eo]; m.dv[en] = d.dvEdge[eo];
        m.fE[en] = d.fEdge ? d.fEdge[eo] : 0.0;  // HorzMesh.jl:257-262
        const int n = eo < nEo ? d.nEdgesOnEdge[eo] : 0;                       // halo rows are never read
        m.nEoE[en] = (uint8_t)n;
        for (int i = 0; i < n; ++i) {
            int32_t x = d.edgesOnEdge[(int64_t)S2f * eo + i];
            m.eoe[(size_t)i * nE + en] = x ? invE[x - 1] : -1;  // 0 entries are skipped (coriolis kernel :67)
            // absent slots (0 entries) carry weight 0 like the padding: every kernel -- fused ForwardEuler reads the raw weights
            // with the padded index rows -- then agrees with the reference kernel, which skips them (coriolis kernel :67)
            m.woe[(size_t)i * nE + en] = (bnd || x == 0) ? 0.0 : d.weightsOnEdge[(int64_t)S2f * eo + i];
        }
        if (eo >= nEo && en >= nEo) {                                          // (halo edges keep their places behind the owned ones)
            const int nh = std::min<int>(std::max<int>(d.nEdgesOnEdge[eo], 0), S2f);
            for (int i = 0; i < nh; ++i) {
                const int32_t x = d.edgesOnEdge[(int64_t)S2f * eo + i];
                if (x < 1 || x > nE) continue;
                m.haloEoe[(size_t)(en - nEo) * S2f + i] = invE[x - 1];
                m.haloWoe[(size_t)(en - nEo) * S2f + i] = bnd ? 0.0 : d.weightsOnEdge[(int64_t)S2f * eo + i];
            }
        }
    }
    m.f0 = m.fE[0];
    for (int64_t e = 0; e < nE; ++e)
        if (m.fE[e] != m.f0) { m.uniformF = false; break; }
    // ---- cells (edgeSignOnCell derived as HorzMesh.jl:292-311 when not supplied) ---------------------
    m.eoc.assign((size_t)S * nC, -1); m.sgnC.assign((size_t)S * nC, 0);
    m.nEoC.resize(nC); m.area.resize(nC); m.H.resize(nC);
#pragma omp parallel for schedule(static)
    for (int64_t cn = 0; cn < nC; ++cn) {
        const int64_t co = m.permC[cn];
        const int n = co < nCo ? d.nEdgesOnCell[co] : 0;                       // halo rows are never read
        m.nEoC[cn] = (uint8_t)n; m.area[cn] = d.areaCell[co]; m.H[cn] = d.restingThicknessSum[co];
        for (int i = 0; i < n; ++i) {
            int32_t e = d.edgesOnCell[(int64_t)Sf * co + i];
            m.eoc[(size_t)i * nC + cn] = invE[e - 1];
            m.sgnC[(size_t)i * nC + cn] = d.edgeSignOnCell ? d.edgeSignOnCell[(int64_t)Sf * co + i]
                                                           : ((int32_t)(co + 1) == d.cellsOnEdge[2 * (int64_t)(e - 1)] ? -1 : 1);
        }
    }
    // ---- vertices (edgeSignOnVertex derived as HorzMesh.jl:313-332 when not supplied) ----------------
    m.eov.assign((size_t)D * nV, -1); m.sgnV.assign((size_t)D * nV, 0); m.areaTri.resize(nV);
#pragma omp parallel for schedule(static)
    for (int64_t vn = 0; vn < nV; ++vn) {
        const int64_t vo = m.permV[vn];
        m.areaTri[vn] = d.areaTriangle[vo];
        for (int j = 0; j < D; ++j) {
            int32_t e = d.edgesOnVertex[(int64_t)D * vo + j];
            m.eov[(size_t)j * nV + vn] = invE[e - 1];
            m.sgnV[(size_t)j * nV + vn] = d.edgeSignOnVertex ? d.edgeSignOnVertex[(int64_t)d.maxEdges * vo + j]
                                                             : ((int32_t)(vo + 1) == d.verticesOnEdge[2 * (int64_t)(e - 1)] ? -1 : 1);
        }
    }
    // ---- blocks of the fused kernel: block b = owned cells [b*kBlockCells, ...) + the owned edges whose first
    //      cell lies in that range (edges are sorted by it).  A block is "boundary" when any of its stencils reads
    //      a halo cell or halo edge; interior blocks can run while the halo exchange is in flight. ---------------
    const int nb = (int)((nCo + kBlockCells - 1) / kBlockCells);
    m.blkEdgeStart.assign(nb + 1, 0);
    {
        int64_t e = 0;
        for (int b = 0; b <= nb; ++b) {
            const int64_t cfirst = (int64_t)b * kBlockCells;
            while (e < nEo && m.ce[2 * e] < cfirst) ++e;
            m.blkEdgeStart[b] = (int32_t)e;
        }
        m.blkEdgeStart[nb] = (int32_t)nEo;
    }
    for (int b = 0; b < nb; ++b) {
        bool halo = false;
        for (int64_t e = m.blkEdgeStart[b]; e < m.blkEdgeStart[b + 1] && !halo; ++e) {
            if (m.ce[2 * e + 1] >= nCo) halo = true;
            for (int i = 0; i < m.nEoE[e] && !halo; ++i) halo = m.eoe[(size_t)i * nE + e] >= nEo;
        }
        for (int64_t c = (int64_t)b * kBlockCells; c < std::min<int64_t>(nCo, (int64_t)(b + 1) * kBlockCells) && !halo; ++c)
            for (int i = 0; i < m.nEoC[c] && !halo; ++i) {
                const int64_t e = m.eoc[(size_t)i * nC + c];
                halo = e >= nEo || m.ce[2 * e] >= nCo || m.ce[2 * e + 1] >= nCo;
            }
        (halo ? m.blkBoundary : m.blkInterior).push_back(b);
    }
    // ---- derived edgesOnEdge -----------------------------------------------------------------------------------------
    // MPAS builds edgesOnEdge[:, e] as: the other edges of cellsOnEdge[1, e] in edgesOnCell order starting after e,
    // then the same for cellsOnEdge[2, e].  Where that holds, the 4*nEdgesOnEdge index bytes per edge (20 % of the
    // fused stage's HBM traffic in Float64) need not be read: the kernel rebuilds the row from the edgesOnCell rows
    // (L1-resident, the cell phase reads them anyway) plus one byte per edge (its position in the two rows).  Verified edge by edge
    // here, never assumed; a block with any non-conforming edge (or a halo cell, whose rows are not stored) keeps
    // reading the explicit array (MOKAB_MESH_EXPLICIT_EOE forces that everywhere).
    m.posE.assign(nE, 0);
    m.blkDerived.assign(nb, 0);
    // the rebuild lives in the compile-time-width kernels: (10, 6) hexagon meshes and (12, 7) meshes that mix in heptagons
    if (!(flags & MOKAB_MESH_EXPLICIT_EOE) && ((S == 6 && S2 == 10) || (S == 7 && S2 == 12))) {
#pragma omp parallel for schedule(static)
        for (int b = 0; b < nb; ++b) {
            bool ok = true;
            for (int64_t e = m.blkEdgeStart[b]; e < m.blkEdgeStart[b + 1]; ++e) {
                const int32_t c1 = m.ce[2 * e], c2 = m.ce[2 * e + 1];
                const bool masked = c1 == c2;
                const int n1 = m.nEoC[c1], n2 = masked ? 1 : m.nEoC[c2];
                int p1 = -1, p2 = masked ? 0 : -1;
                for (int i = 0; i < n1; ++i) if (m.eoc[(size_t)i * nC + c1] == e) p1 = i;
                if (!masked) for (int i = 0; i < n2; ++i) if (m.eoc[(size_t)i * nC + c2] == e) p2 = i;
                if (c1 / kBlockCells != b || p1 < 0 || p2 < 0 || n1 + n2 - 2 != m.nEoE[e]) { ok = false; continue; }
                m.posE[e] = (uint8_t)(p1 | (p2 << 3) | ((!masked && n1 == S && n2 == S) ? 128 : 0));
                for (int j = 0; j < m.nEoE[e] && ok; ++j) {
                    int32_t want;
                    if (j < n1 - 1) { int r = p1 + 1 + j; r -= r >= n1 ? n1 : 0; want = m.eoc[(size_t)r * nC + c1]; }
                    else { int r = p2 + 1 + (j - (n1 - 1)); r -= r >= n2 ? n2 : 0; want = m.eoc[(size_t)r * nC + c2]; }
                    ok = m.eoe[(size_t)j * nE + e] == want;
                }
            }
            m.blkDerived[b] = ok ? 1 : 0;
        }
    }
}

// Arrays of the fused RK4 path in precision R: f folded into the weights, g/dc and 1/area
// precomputed, the cell-side sign carried in bit 0 of the edge id.
template <class R>
struct FusedMesh {
    DevBuf<R> gdc, wf, dv, invArea, H;
    DevBuf<R> wfT;             // adjoint: transposed Coriolis weights (see moka_b200.cu: ensure_adjoint_mesh)
    DevBuf<R> wfI;             // TMA = 3 stage variant: the weights slot-interleaved, 16 bytes per edge and slot group (built on first use)
    DevBuf<R> wfB;             // TMA = 2 stage variant: the weights block-major (built on first use)
    DevBuf<long long> wfBOff;  //                        first element of every block's run
    bool ready = false;
};

}  // namespace mokab

struct mokab_mesh {
    mokab_ctx *ctx = nullptr;
    int64_t nC = 0, nE = 0, nV = 0;
    int S = 0, S2 = 0, D = 0;
    std::vector<int32_t> permC, permE, permV;
    // reference-form arrays (Float64, unfolded) for the operator-level kernels
    mokab::DevBuf<int2> ce;
    mokab::DevBuf<int32_t> eoe, eoc, sgnC, eov, sgnV, dPermC, dPermE, dPermV;
    mokab::DevBuf<uint8_t> nEoE, nEoC;
    mokab::DevBuf<double> woe, dc, dv, fE, area, H, areaTri;
    // fused-path arrays
    mokab::DevBuf<int32_t> eoeF, eocF;   // absent -> self / sign in bit 0
    mokab::DevBuf<int32_t> blkEdgeStart; // per block of kBlockCells cells: first owned edge
    mokab::DevBuf<int32_t> blkInterior, blkBoundary;  // block ids by part
    mokab::DevBuf<uint8_t> posE, blkDerived;          // derived edgesOnEdge (mesh.cuh: build_host_mesh)
    int nDerivedBlocks = 0;
    int64_t nCo = 0, nEo = 0;            // owned cells / edges (== nC / nE without decomposition)
    int fusedBlocks = 0, nInterior = 0, nBoundary = 0;
    int maxBlockEdges = 0;               // the most edges any block owns (sizes the shared-memory rows of the TMA stage variant)
    bool uniformF = true; double f0 = 0.0;
    std::vector<int32_t> hBlkEdgeStart, hBlkInterior, hBlkBoundary;  // host copies (halo_setup re-classifies)
    mokab::DevBuf<int32_t> haloSend, haloRecv;  // combined [cells | edges] indices, device numbering
    std::vector<int32_t> hHaloSend;             // host copy (the push tables of the direct-store exchange are built from it)
    bool halo_ready = false;
    std::vector<int32_t> hHaloEoe;              // decomposed meshes: the halo edges' rows (HostMesh::haloEoe), for the reverse mode
    std::vector<double> hHaloWoe;
    int haloS2 = 0;
    // adjoint: transpose of the Coriolis stencil (built on first use)
    mokab::DevBuf<int32_t> eoeT;
    mokab::DevBuf<double> woeT;
    mokab::DevBuf<uint8_t> nEoET;
    int S2T = 0;
    bool adj_ready = false;
    mokab::FusedMesh<double> f64;
    mokab::FusedMesh<float> f32;
    int64_t device_bytes() const
    {
        return (int64_t)(ce.bytes() + eoe.bytes() + eoc.bytes() + sgnC.bytes() + eov.bytes() + sgnV.bytes() + dPermC.bytes() +
                         dPermE.bytes() + dPermV.bytes() + nEoE.bytes() + nEoC.bytes() + woe.bytes() + dc.bytes() + dv.bytes() +
                         fE.bytes() + area.bytes() + H.bytes() + areaTri.bytes() + eoeF.bytes() + eocF.bytes() +
                         blkEdgeStart.bytes() + f64.gdc.bytes() + f64.wf.bytes() + f64.dv.bytes() + f64.invArea.bytes() +
                         f64.H.bytes() + f32.gdc.bytes() + f32.wf.bytes() + f32.dv.bytes() + f32.invArea.bytes() + f32.H.bytes());
    }
};
