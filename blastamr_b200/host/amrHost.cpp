// amrHost.cpp -- see amrHost.H.  Every numerical loop of the reference is ONE call into libamr_b200.so.
#include "amrHost.H"

#include <algorithm>
#include <cstring>

namespace amr {

// ------------------------------------------------------------------------------------------------ fvMeshLite

label fvMeshLite::findPatchID(const word &nm) const {
    for (size_t i = 0; i < patches.size(); i++) if (patches[i].name == nm) return (label)i;
    return -1;
}
volField &fvMeshLite::lookupField(const word &nm) {
    auto it = fields.find(nm);
    if (it == fields.end()) throw FatalError("request for volScalarField " + nm + " from objectRegistry failed");
    return it->second;
}

// mag(field) for non-scalar fields: errorEstimator::getFieldValueType, errorEstimatorTemplates.C:30-46
static scalarField magField(const volField &f, int64_t n) {
    if ((int64_t)f.v.size() != n * f.nComp) throw FatalError("field size does not match the mesh");
    if (f.nComp == 1) return f.v;
    scalarField m((size_t)n);
    for (int64_t c = 0; c < n; c++) {
        scalar s = 0;
        for (int k = 0; k < f.nComp; k++) s += f.v[c * f.nComp + k] * f.v[c * f.nComp + k];
        m[c] = std::sqrt(s);
    }
    return m;
}

// ------------------------------------------------------------------------------------------------ errorEstimator

std::map<word, errorEstimator::ctor> &errorEstimator::dictionaryConstructorTable() {
    static std::map<word, ctor> t;
    return t;
}

std::unique_ptr<errorEstimator> errorEstimator::New(fvMeshLite &mesh, amr_ctx *gpu, const dictionary &dict) {
    const word type = dict.lookupWord("errorEstimator");          // errorEstimatorNew.C:36
    auto &t = dictionaryConstructorTable();
    auto it = t.find(type);
    if (it == t.end()) {
        std::string valid;
        for (auto &e : t) valid += e.first + " ";
        throw FatalError("Unknown errorEstimator type " + type + "\n\nValid errorEstimator types :\n" + valid);
    }
    return it->second(mesh, gpu, dict, word());
}

errorEstimator::errorEstimator(fvMeshLite &mesh, amr_ctx *gpu, const dictionary &dict, const word &name)
    : mesh_(mesh), gpu_(gpu), name_(name),
      protectedPatches_(dict.lookupOrDefault("protectedPatches", wordList{})),       // errorEstimator.C:145
      nPatchesBuffers_((label)dict.lookupOrDefaultLabel("nPatchesBuffers", 1)),      // :146
      nBufferCells_((label)std::max(dict.lookupOrDefaultLabel("nBufferLayers", 1),
                                    dict.lookupOrDefaultLabel("nRefinementBufferLayers", 1))),   // :147-154
      maxRefinementLevel_((label)dict.lookupOrDefaultLabel("maxRefinement", 1)) {}   // :155

void errorEstimator::check(int rc, const char *where) const {
    if (rc != AMR_OK) throw FatalError(std::string(where) + ": " + amr_last_error(gpu_));
}

void errorEstimator::read(const dictionary &dict) {          // errorEstimator.C:167-191
    lowerRefine_ = dict.lookupScalar("lowerRefineLevel");
    lowerUnrefine_ = dict.lookupScalar("unrefineLevel");
    upperRefine_ = dict.lookupOrDefault("upperRefineLevel", GREAT);
    upperUnrefine_ = dict.lookupOrDefault("upperUnrefineLevel", GREAT);
    if (dict.found("maxRefinement")) { maxLevel_ = (label)dict.lookupLabel("maxRefinement"); minDx_ = -1; }     // :175-190
    else if (dict.found("minDx")) { minDx_ = dict.lookupScalar("minDx"); maxLevel_ = -1; }
    else throw FatalError("Either maxRefinement or minDx must be specified");
    protectedPatches_ = dict.lookupOrDefault("protectedPatches", wordList{});
}

void errorEstimator::geometricD(int32_t gd[3]) const {
    gd[0] = 1; gd[1] = 1; gd[2] = mesh_.nSolutionD == 2 ? -1 : 1;      // the empty direction of the 2-D cases is z
}
// errorEstimator.C:262-292: uniform maxRefinement, or cellLevel + 1 where the cell is still wider than minDx and flagged
labelList errorEstimator::maxRefinement() const {
    labelList out((size_t)mesh_.nCells, maxLevel_);
    if (maxLevel_ >= 0 || mesh_.cellLevel.empty()) return out;
    if (mesh_.Sf.empty()) throw FatalError("errorEstimator::maxRefinement: minDx needs the mesh geometry (meshSizeObject)");
    int32_t gd[3];
    geometricD(gd);
    scalarField dx((size_t)mesh_.nCells);
    check(amr_mesh_size(gpu_, gd, dx.data(), nullptr), "meshSizeObject::dx");
    check(amr_max_refinement(gpu_, -1, minDx_, dx.data(), error_.data(), out.data()), "errorEstimator::maxRefinement");
    return out;
}

void errorEstimator::normalize(scalarField &error) {
    check(amr_normalize(gpu_, lowerRefine_, upperRefine_, lowerUnrefine_, upperUnrefine_, error.data()), "errorEstimator::normalize");
}

label errorEstimator::protectPatches() {
    if (protectedPatches_.empty()) return 0;                  // errorEstimator.C:331-334
    labelList ids;
    for (const word &nm : protectedPatches_) {
        label id = mesh_.findPatchID(nm);
        if (id >= 0) ids.push_back(id);
    }
    int64_t nProt = 0;
    const int nLayers = nPatchesBuffers_ + (maxRefinementLevel_ - 1) * nBufferCells_;   // :347
    check(amr_protect_patches(gpu_, ids.data(), (int)ids.size(), nLayers, error_.data(), &nProt), "errorEstimator::protectPatches");
    return (label)nProt;
}

namespace errorEstimators {

#define THR4 const double thr[4] = {lowerRefine_, upperRefine_, lowerUnrefine_, upperUnrefine_}

struct delta : errorEstimator {           // delta.C
    word fieldName_;
    delta(fvMeshLite &m, amr_ctx *g, const dictionary &d, const word &n) : errorEstimator(m, g, d, n), fieldName_(d.lookupWord("deltaField")) { read(d); }
    word type() const override { return "delta"; }
    void update(const bool scale) override {
        const volField &x = mesh_.lookupField(fieldName_);
        if (x.nComp != 1) throw FatalError("delta: " + fieldName_ + " is not a volScalarField");
        error_.resize((size_t)mesh_.nCells);
        THR4;
        check(amr_estimate(gpu_, AMR_EST_DELTA, x.v.data(), nullptr, 0, 0, scale, thr, error_.data()), "delta::update");
    }
};
struct scaledDelta : errorEstimator {     // scaledDelta.C
    word fieldName_;
    scalar minVal_, offset_;
    scaledDelta(fvMeshLite &m, amr_ctx *g, const dictionary &d, const word &n)
        : errorEstimator(m, g, d, n), fieldName_(d.lookupWord("scaledDeltaField")), minVal_(d.lookupOrDefault("minValue", SMALL)),
          offset_(d.lookupOrDefault("offset", 0.0)) { read(d); }
    word type() const override { return "scaledDelta"; }
    void update(const bool scale) override {
        const scalarField x = magField(mesh_.lookupField(fieldName_), mesh_.nCells);
        error_.resize((size_t)mesh_.nCells);
        THR4;
        check(amr_estimate(gpu_, AMR_EST_SCALED_DELTA, x.data(), nullptr, minVal_, offset_, scale, thr, error_.data()), "scaledDelta::update");
    }
};
struct densityGradient : errorEstimator { // absent from the reference; see include/amr_b200.h AMR_EST_DENSITY_GRADIENT
    word fieldName_;
    scalar minVal_;
    densityGradient(fvMeshLite &m, amr_ctx *g, const dictionary &d, const word &n)
        : errorEstimator(m, g, d, n), fieldName_(d.lookupOrDefault("densityField", word("rho"))), minVal_(d.lookupOrDefault("minValue", SMALL)) { read(d); }
    word type() const override { return "densityGradient"; }
    void update(const bool scale) override {
        const scalarField x = magField(mesh_.lookupField(fieldName_), mesh_.nCells);
        error_.resize((size_t)mesh_.nCells);
        THR4;
        check(amr_estimate(gpu_, AMR_EST_DENSITY_GRADIENT, x.data(), nullptr, minVal_, 0.0, scale, thr, error_.data()), "densityGradient::update");
    }
};
struct Lohner : errorEstimator {          // Lohner.C
    word fieldName_;
    scalar epsilon_;
    Lohner(fvMeshLite &m, amr_ctx *g, const dictionary &d, const word &n)
        : errorEstimator(m, g, d, n), fieldName_(d.lookupWord("deltaField")), epsilon_(d.lookupScalar("epsilon")) { read(d); }
    word type() const override { return "Lohner"; }
    void update(const bool scale) override {
        const volField &x = mesh_.lookupField(fieldName_);
        if (x.nComp != 1) throw FatalError("Lohner: " + fieldName_ + " is not a volScalarField");
        error_.resize((size_t)mesh_.nCells);
        THR4;
        check(amr_estimate_lohner(gpu_, x.v.data(), nullptr, nullptr, epsilon_, scale, thr, error_.data()), "Lohner::update");
    }
};
struct fieldValue : errorEstimator {      // fieldValue.C
    word fieldName_;
    fieldValue(fvMeshLite &m, amr_ctx *g, const dictionary &d, const word &n) : errorEstimator(m, g, d, n), fieldName_(d.lookupWord("fieldName")) { read(d); }
    word type() const override { return "fieldValue"; }
    void update(const bool scale) override {
        const scalarField x = magField(mesh_.lookupField(fieldName_), mesh_.nCells);
        error_.resize((size_t)mesh_.nCells);
        THR4;
        check(amr_estimate(gpu_, AMR_EST_FIELD_VALUE, x.data(), nullptr, 0, 0, scale, thr, error_.data()), "fieldValue::update");
    }
};
struct gradientRange : errorEstimator {   // gradientRange.C ("gradientRange" in the run-time table)
    word gradField_;
    scalar refineThreshold_, unrefineThreshold_;
    gradientRange(fvMeshLite &m, amr_ctx *g, const dictionary &d, const word &n)
        : errorEstimator(m, g, d, n), gradField_(d.lookupWord("gradField")), refineThreshold_(d.lookupOrDefault("refineThreshold", 0.5)),
          unrefineThreshold_(d.lookupOrDefault("unrefineThreshold", 0.4)) { read(d); }
    word type() const override { return "gradientRange"; }
    void read(const dictionary &d) override {
        errorEstimator::read(d);
        refineThreshold_ = d.lookupOrDefault("refineThreshold", 0.5);
        unrefineThreshold_ = d.lookupOrDefault("unrefineThreshold", 0.4);
    }
    void update(const bool scale) override {
        double thr[4];
        const word cached = "mag(grad(" + gradField_ + "))";
        if (mesh_.fields.count(cached)) {
            // a solver that already holds fvc::grad (any scheme) hands mag(grad(f)) over
            error_ = magField(mesh_.lookupField(cached), mesh_.nCells);
            check(amr_normalize_range(gpu_, refineThreshold_, unrefineThreshold_, scale, error_.data(), thr, nullptr), "gradientEstimator::update");
        } else {
            // fvc::grad with the case's gradScheme on the device (gradientRange.C:100)
            const volField &x = mesh_.lookupField(gradField_);
            if (x.nComp != 1) throw FatalError("gradientRange: " + gradField_ + " is not a volScalarField");
            if (mesh_.Sf.empty()) throw FatalError("gradientRange: fvc::grad needs the mesh geometry");
            int scheme = AMR_GRAD_GAUSS_LINEAR;
            if (mesh_.gradScheme == "leastSquares") {
                scheme = AMR_GRAD_LEAST_SQUARES;
                if (mesh_.lsPVectors.empty()) throw FatalError("gradientRange: leastSquares needs leastSquaresVectors");
                check(amr_mesh_set_ls_vectors(gpu_, mesh_.lsPVectors.data(), mesh_.lsNVectors.data()), "leastSquaresVectors");
            } else if (mesh_.gradScheme != "Gauss linear") throw FatalError("gradientRange: gradScheme " + mesh_.gradScheme + " is not on the device path");
            error_.resize((size_t)mesh_.nCells);
            check(amr_estimate_gradient(gpu_, scheme, x.v.data(), nullptr, nullptr, 0.0, refineThreshold_, unrefineThreshold_, scale, thr,
                                        nullptr, error_.data()), "gradientEstimator::update");
        }
        lowerRefine_ = thr[0]; upperRefine_ = thr[1]; lowerUnrefine_ = thr[2]; upperUnrefine_ = thr[3];
    }
};
struct cellSize : errorEstimator {        // cellSize.C
    int sizeType_ = AMR_SIZE_CHARACTERISTIC;
    labelList cmpts_;
    double minDX_[3] = {0, 0, 0}, maxDX_[3] = {0, 0, 0};
    cellSize(fvMeshLite &m, amr_ctx *g, const dictionary &d, const word &n) : errorEstimator(m, g, d, n) { read(d); }
    word type() const override { return "cellSize"; }
    static void readVector(const dictionary &d, const word &key, double v[3]) {
        const wordList w = d.lookupWordList(key);
        if (w.size() != 3) throw FatalError("keyword " + key + ": expected a vector");
        for (int k = 0; k < 3; k++) v[k] = std::stod(w[(size_t)k]);
    }
    void read(const dictionary &d) override {                 // cellSize.C:213-232 (does not call the base read)
        const word t = d.lookupWord("type");
        if (t == "volume") sizeType_ = AMR_SIZE_VOLUME;
        else if (t == "characteristic") sizeType_ = AMR_SIZE_CHARACTERISTIC;
        else if (t == "mag") sizeType_ = AMR_SIZE_MAG;
        else if (t == "cmpt") sizeType_ = AMR_SIZE_CMPT;
        else throw FatalError("cellSize: unknown size type " + t + "; valid types: volume cmpt characteristic mag");     // cellSize.C:42-48
        if (sizeType_ == AMR_SIZE_CMPT) {
            int32_t gd[3];
            geometricD(gd);
            cmpts_.clear();
            for (const word &c : d.lookupWordList("cmpts")) {           // readCmpts, cellSize.C:76-121
                label k;
                if (c == "x" || c == "X" || c == "0") k = 0;
                else if (c == "y" || c == "Y" || c == "1") k = 1;
                else if (c == "z" || c == "Z" || c == "2") k = 2;
                else throw FatalError("Unknown component " + c + "\nValid components are: \nx/X\ny/Y\nz/Z");
                if (gd[k] == 1 && std::find(cmpts_.begin(), cmpts_.end(), k) == cmpts_.end()) cmpts_.push_back(k);
            }
            readVector(d, "minDX", minDX_);
            readVector(d, "maxDX", maxDX_);
        } else if (sizeType_ == AMR_SIZE_VOLUME) {
            lowerUnrefine_ = d.lookupScalar("minVolume");
            lowerRefine_ = d.lookupScalar("maxVolume");
        } else {
            lowerUnrefine_ = d.lookupScalar("minDx");
            lowerRefine_ = d.lookupScalar("maxDx");
        }
        maxLevel_ = (label)d.lookupOrDefaultLabel("maxRefinement", 10);
    }
    labelList maxRefinement() const override {                // cellSize.C:124-127
        labelList out(mesh_.cellLevel);
        for (label &l : out) l++;
        return out;
    }
    void update(const bool) override {
        if (mesh_.Sf.empty()) throw FatalError("cellSize: needs the mesh geometry (meshSizeObject)");
        int32_t gd[3];
        geometricD(gd);
        const size_t n = (size_t)mesh_.nCells;
        scalarField dx, dX;
        if (sizeType_ == AMR_SIZE_CHARACTERISTIC) dx.resize(n);
        if (sizeType_ == AMR_SIZE_MAG || sizeType_ == AMR_SIZE_CMPT) dX.resize(3 * n);
        if (!dx.empty() || !dX.empty())
            check(amr_mesh_size(gpu_, gd, dx.empty() ? nullptr : dx.data(), dX.empty() ? nullptr : dX.data()), "meshSizeObject");
        error_.resize(n, 0.0);                                 // CMPT accumulates into the field it finds
        check(amr_estimate_cell_size(gpu_, sizeType_, gd, lowerUnrefine_, lowerRefine_, cmpts_.data(), (int)cmpts_.size(), minDX_, maxDX_,
                                     dx.empty() ? nullptr : dx.data(), dX.empty() ? nullptr : dX.data(), error_.data()), "cellSize::update");
    }
};
struct multicomponent : errorEstimator {  // multicomponentError/multicomponent.C
    wordList names_;
    std::vector<std::unique_ptr<errorEstimator>> errors_;
    multicomponent(fvMeshLite &m, amr_ctx *g, const dictionary &d, const word &n) : errorEstimator(m, g, d, n) {
        for (auto &e : d.lookupEntryList("errorEstimators")) {            // :52-71
            names_.push_back(e.first);
            const word type = e.second.lookupWord("errorEstimator");
            auto it = dictionaryConstructorTable().find(type);
            if (it == dictionaryConstructorTable().end()) throw FatalError("Unknown errorEstimator type " + type);
            errors_.push_back(it->second(m, g, e.second, e.first));
        }
        maxLevel_ = 0;
        for (auto &e : errors_) maxLevel_ = std::max(maxLevel_, e->maxLevel());
    }
    word type() const override { return "multicomponent"; }
    void read(const dictionary &d) override {                 // :84-96
        auto entries = d.lookupEntryList("errorEstimators");
        for (size_t i = 0; i < errors_.size(); i++) errors_[i]->read(entries[i].second);
        maxLevel_ = 0;
        for (auto &e : errors_) maxLevel_ = std::max(maxLevel_, e->maxLevel());
    }
    void update(const bool scale) override {                  // :99-119
        std::vector<const double *> ptr;
        for (auto &e : errors_) { e->update(scale); ptr.push_back(e->error().data()); }
        error_.resize((size_t)mesh_.nCells);
        check(amr_multicomponent(gpu_, (int)ptr.size(), ptr.data(), nullptr, error_.data(), nullptr, nullptr), "multicomponent::update");
    }
    labelList maxRefinement() const override {                // :122-157
        std::vector<const double *> ptr;
        std::vector<labelList> lv;
        std::vector<const int32_t *> lptr;
        for (auto &e : errors_) { ptr.push_back(e->error().data()); lv.push_back(e->maxRefinement()); }
        for (auto &l : lv) lptr.push_back(l.data());
        labelList out((size_t)mesh_.nCells, 0);
        scalarField tmp((size_t)mesh_.nCells);
        check(amr_multicomponent(gpu_, (int)ptr.size(), ptr.data(), lptr.data(), tmp.data(), out.data(), nullptr), "multicomponent::maxRefinement");
        return out;
    }
};

template <class T>
static void add(const word &nm) {
    errorEstimator::dictionaryConstructorTable()[nm] = [](fvMeshLite &m, amr_ctx *g, const dictionary &d, const word &n) {
        return std::unique_ptr<errorEstimator>(new T(m, g, d, n));
    };
}
static bool registered = []() {
    add<delta>("delta"); add<scaledDelta>("scaledDelta"); add<Lohner>("Lohner"); add<fieldValue>("fieldValue");
    add<gradientRange>("gradientRange"); add<densityGradient>("densityGradient");
    add<cellSize>("cellSize"); add<multicomponent>("multicomponent");
    errorEstimator::dictionaryConstructorTable()["coded"] = [](fvMeshLite &, amr_ctx *, const dictionary &, const word &) -> std::unique_ptr<errorEstimator> {
        throw FatalError("coded errorEstimator: user C++ is compiled and run by the host (codedErrorEstimator.C); "
                         "its error_ field enters the device path through amr_normalize / amr_select_refine");
    };
    return true;
}();

}  // namespace errorEstimators

// ------------------------------------------------------------------------------------------------ fvMeshRefiner

std::map<word, fvMeshRefiner::ctor> &fvMeshRefiner::dictionaryConstructorTable() {
    static std::map<word, ctor> t;
    return t;
}
std::unique_ptr<fvMeshRefiner> fvMeshRefiner::New(adaptiveFvMesh &mesh, const dictionary &dict) {
    word type = "hexRefiner";                                   // fvMeshRefinerNew.C:82-86
    if (dict.found("refiner")) type = dict.lookupWord("refiner");
    auto &t = dictionaryConstructorTable();
    auto it = t.find(type);
    if (it == t.end()) {
        std::string valid;
        for (auto &e : t) valid += e.first + " ";
        throw FatalError("Unknown fvMeshRefiner type " + type + "\n\nValid fvMeshRefiner types :\n" + valid);
    }
    return it->second(mesh, dict);
}
static bool refinersRegistered = []() {
    fvMeshRefiner::dictionaryConstructorTable()["hexRefiner"] = [](adaptiveFvMesh &m, const dictionary &d) {
        return std::unique_ptr<fvMeshRefiner>(new fvMeshHexRefiner(m, d));
    };
    auto notYet = [](const char *nm) {
        return [nm](adaptiveFvMesh &, const dictionary &) -> std::unique_ptr<fvMeshRefiner> {
            throw FatalError(std::string(nm) + ": this refiner's edge-consistency sweeps are not on the device path yet (SURVEY.md 8a B2/B4)");
        };
    };
    fvMeshRefiner::dictionaryConstructorTable()["polyRefiner"] = [](adaptiveFvMesh &m, const dictionary &d) {
        return std::unique_ptr<fvMeshRefiner>(new fvMeshPolyRefiner(m, d));
    };
    fvMeshRefiner::dictionaryConstructorTable()["directionalRefiner"] = notYet("directionalRefiner");
    return true;
}();

fvMeshRefiner::fvMeshRefiner(adaptiveFvMesh &mesh, const dictionary &dict) : amesh_(mesh), dict_(dict) { readDict(dict); }

void fvMeshRefiner::readDict(const dictionary &dict) {          // fvMeshRefiner.C:526-613
    dict_ = dict;
    maxCells_ = (label)dict_.lookupOrDefaultLabel("maxCells", labelMax);
    if (maxCells_ <= 0) throw FatalError("Illegal maximum number of cells; the maxCells should be > 0.");
    if (dict.found("nRefinementBufferLayers")) nRefinementBufferLayers_ = (label)dict_.lookupLabel("nRefinementBufferLayers");
    else if (dict.found("nBufferLayers")) nRefinementBufferLayers_ = (label)dict_.lookupLabel("nBufferLayers");
    if (dict.found("nUnrefinementBufferLayers")) nUnrefinementBufferLayers_ = (label)dict_.lookupLabel("nUnrefinementBufferLayers");
    else if (dict.found("nBufferLayers")) nUnrefinementBufferLayers_ = (label)dict_.lookupLabel("nBufferLayers");
    protectedPatchNames_ = dict_.lookupOrDefault("protectedPatches", wordList());
    refine_ = dict_.lookupOrDefaultBool("refine", true);
    unrefine_ = dict_.lookupOrDefaultBool("unrefine", true);
    refineInterval_ = (label)dict_.lookupOrDefaultLabel("refineInterval", 1);
    if (refineInterval_ < 0) throw FatalError("Illegal refineInterval; the refineInterval should be >= 1.");
    unrefineInterval_ = (label)dict_.lookupOrDefaultLabel("unrefineInterval", refineInterval_);
    if (unrefineInterval_ < 0) throw FatalError("Illegal unrefineInterval; the unrefineInterval should be >= 1.");
    beginRefine_ = dict_.lookupOrDefault("beginRefine", 0.0);
    beginUnrefine_ = dict_.lookupOrDefault("beginUnrefine", 0.0);
    endRefine_ = dict_.lookupOrDefault("endRefine", GREAT);
    endUnrefine_ = dict_.lookupOrDefault("endUnrefine", GREAT);
}

bool fvMeshRefiner::canRefine(const bool incr) const {           // fvMeshRefiner.C:274-310
    const fvMeshLite &m = amesh_.mesh();
    if (m.timeIndex <= 1) return true;          // force refinement on first timestep (before `refine` is looked at)
    if (!refine_) return false;
    if (force_) {
    } else if (m.timeValue < beginRefine_ || m.timeValue > endRefine_) return false;
    else if (refineInterval_ > 0 && (m.timeIndex % refineInterval_) > 0) return false;
    if (incr) nRefinementIterations_++;
    // mesh_.globalData().nTotalCells() (fvMeshRefiner.C:308): all ranks must take the same branch, or only some of them
    // would enter the collective amr_select_refine.  amr_imbalance carries the reference's
    // returnReduce(mesh_.nCells(), sumOp<label>()) (fvMeshBalance.C:463); without a communicator it is the local count.
    double ratio = 0;
    int64_t nTotalCells = m.nCells;
    amr_ctx *gpu = const_cast<adaptiveFvMesh &>(amesh_).gpu();
    if (gpu && amr_imbalance(gpu, &ratio, &nTotalCells) != AMR_OK) nTotalCells = m.nCells;
    return nTotalCells < maxCells_;
}
bool fvMeshRefiner::canUnrefine(const bool incr) const {         // fvMeshRefiner.C:313-341
    const fvMeshLite &m = amesh_.mesh();
    if (!unrefine_ || m.timeIndex <= 0) return false;
    if (force_) {
    } else if (m.timeIndex <= 0 || m.timeValue < beginUnrefine_ || m.timeValue > endUnrefine_) return false;
    else if (unrefineInterval_ > 0 && (m.timeIndex % unrefineInterval_) > 0) return false;
    if (incr) nUnrefinementIterations_++;
    return true;
}

// fvMeshHexRefiner::refine, fvMeshHexRefiner.C:1476-1684
bool fvMeshHexRefiner::refine(const scalarField &errorIn, const labelList &maxCellLevel, const scalar lowerRefineLevel,
                              const scalar upperRefineLevel, const scalar unrefineLevel) {
    readDict(this->dict_);
    bool hasChanged = false;
    fvMeshLite &mesh = amesh_.mesh();
    amr_ctx *gpu = amesh_.gpu();
    lastCellsToRefine.clear(); lastElemsToUnrefine.clear();
    if (!(canRefine() || canUnrefine())) return false;            // preUpdate(), :1487
    scalarField error(errorIn);
    boolList refineCell((size_t)mesh.nCells, 0);
    labelList pp;
    for (const word &nm : protectedPatchNames_) { label id = mesh.findPatchID(nm); if (id >= 0) pp.push_back(id); }
    const int kind = amesh_.kindOverride >= 0 ? amesh_.kindOverride : amesh_.refinerKind();

    if (canRefine(true)) {
        labelList maxRefinement(maxCellLevel);
        if (maxRefinement.empty()) {                              // setMaxCellLevel, fvMeshRefiner.C:222-253
            if (!dict_.found("maxRefinement")) throw FatalError("maxRefinement was not specified");
            maxRefinement.assign((size_t)mesh.nCells, (label)dict_.lookupLabel("maxRefinement"));
        }
        if (!maxRefinement.empty() && *std::min_element(maxRefinement.begin(), maxRefinement.end()) < 0)
            throw FatalError("Illegal maximum refinement level; the maxRefinement should be > 0.");
        if ((int64_t)maxRefinement.size() != mesh.nCells) throw FatalError("Inconsistent number of cells and size of maxCellsLevel");
        labelList cellsToRefine((size_t)std::max<int64_t>(mesh.nCells, 1));
        int64_t nCellsToRefine = 0;
        amesh_.check(amr_select_refine(gpu, error.data(), lowerRefineLevel, upperRefineLevel, maxRefinement.data(), 0,
                                       nRefinementBufferLayers_, pp.data(), (int)pp.size(), maxCells_, kind, refineCell.data(),
                                       cellsToRefine.data(), &nCellsToRefine, nullptr),
                     "fvMeshHexRefiner::refine (selectRefineCells)");
        cellsToRefine.resize((size_t)nCellsToRefine);
        lastCellsToRefine = cellsToRefine;
        if (nCellsToRefine > 0) {
            if (!amesh_.topo()) throw FatalError("no topology changer attached (polyTopoChange is host bookkeeping)");
            const int64_t nOld = mesh.nCells;
            mapPolyMesh map = amesh_.topo()->refine(mesh, cellsToRefine);     // refine(cellsToRefine), :183-246
            // refineCell through the map, :1562-1586 (before the new mesh is uploaded: the marking survives amr_mesh_set)
            boolList newRefineCell(map.cellMap.size());
            amesh_.check(amr_remap_refine_cell(gpu, nOld, (int64_t)map.cellMap.size(), map.cellMap.data(), map.reverseCellMap.data(),
                                               refineCell.data(), newRefineCell.data()),
                         "fvMeshHexRefiner::refine (refineCell remap)");
            refineCell.swap(newRefineCell);
            // the error field is a registered volScalarField: mapped with the others
            scalarField newError(map.cellMap.size());
            amesh_.check(amr_map_field(gpu, nOld, (int64_t)map.cellMap.size(), map.cellMap.data(), 1, error.data(), newError.data(),
                                       AMR_MAP_DIRECT, nullptr, nullptr), "fvMesh::mapFields(error)");
            error.swap(newError);
            amesh_.updateMesh(map);                                            // mesh_.updateMesh(map) -> mapFields
            int64_t nBad = 0;
            amesh_.check(amr_check_levels(gpu, nullptr, &nBad), "hexRef::checkRefinementLevels");   // :243
            hasChanged = true;
        }
    }

    if (canUnrefine(true)) {
        const int64_t nElemMax = std::max<int64_t>(std::max(mesh.nPoints, mesh.nEdges), 1);
        labelList elems((size_t)nElemMax);
        int64_t nSplitElems = 0;
        amesh_.check(amr_select_unrefine(gpu, error.data(), unrefineLevel, refineCell.data(), nUnrefinementBufferLayers_, pp.data(),
                                         (int)pp.size(), kind, elems.data(), &nSplitElems, nullptr),
                     "fvMeshHexRefiner::refine (selectUnrefineElems)");
        elems.resize((size_t)nSplitElems);
        lastElemsToUnrefine = elems;
        if (nSplitElems > 0) {
            if (!amesh_.topo()) throw FatalError("no topology changer attached");
            mapPolyMesh map = amesh_.topo()->unrefine(mesh, elems);           // unrefine(elems), :249-303
            amesh_.updateMesh(map);
            int64_t nBad = 0;
            amesh_.check(amr_check_levels(gpu, nullptr, &nBad), "hexRef::checkRefinementLevels");   // :300
            hasChanged = true;
        }
    }
    return hasChanged;
}

// fvMeshPolyRefiner::readDict, fvMeshPolyRefiner.C:260-301
void fvMeshPolyRefiner::readDict(const dictionary &dict) {
    fvMeshRefiner::readDict(dict);
    if (nRefinementBufferLayers_ < 3 && amesh_.mesh().nSolutionD == 2) nRefinementBufferLayers_ = 3;   // :267-283
    if (nUnrefinementBufferLayers_ < 2) nUnrefinementBufferLayers_ = 2;                                // :287-300
}
// fvMeshPolyRefiner::refine, fvMeshPolyRefiner.C:304-489: same driver shape as the hex refiner; the cutter-specific
// parts (edge consistency, point buffer layers, pointLevel split points) live behind AMR_REFINER_POLY3D, or
// AMR_REFINER_POLY2D when the constructor would have picked prismatic2DRefinement (nGeometricD == 2, :139-155)
bool fvMeshPolyRefiner::refine(const scalarField &error, const labelList &maxCellLevel, const scalar lo, const scalar hi,
                               const scalar unref) {
    amesh_.check(amr_set_option(amesh_.gpu(), AMR_OPT_EDGE_BASED_CONSISTENCY, dict_.lookupOrDefaultBool("edgeBasedConsistency", true)),
                 "refinement::edgeBasedConsistency");
    amesh_.check(amr_set_option(amesh_.gpu(), AMR_OPT_POLY_FORCE, force_ ? 1 : 0), "fvMeshPolyRefiner::force_");
    amesh_.kindOverride = amesh_.mesh().nSolutionD == 2 ? AMR_REFINER_POLY2D : AMR_REFINER_POLY3D;
    return fvMeshHexRefiner::refine(error, maxCellLevel, lo, hi, unref);
}

// ------------------------------------------------------------------------------------------------ adaptiveFvMesh

adaptiveFvMesh::adaptiveFvMesh(const std::string &text, int device) {
    if (amr_ctx_create(device, &gpu_) != AMR_OK)
        throw FatalError("adaptiveFvMesh: no usable CUDA device (this library has no CPU path)");
    setDict(text);
}
adaptiveFvMesh::~adaptiveFvMesh() {
    error_.reset(); refiner_.reset();
    if (gpu_) amr_ctx_destroy(gpu_);
}
void adaptiveFvMesh::check(int rc, const char *where) const {
    if (rc != AMR_OK) throw FatalError(std::string(where) + ": " + amr_last_error(gpu_));
}
void adaptiveFvMesh::setDict(const std::string &text) {
    dictText_ = text;
    dict_ = dictionary::parse(text);
    if (dict_.found("dynamicFvMesh") && dict_.lookupWord("dynamicFvMesh") != "adaptiveFvMesh")
        throw FatalError("Unknown dynamicFvMesh type " + dict_.lookupWord("dynamicFvMesh") + "\n\nValid dynamicFvMesh types :\nadaptiveFvMesh");
}
void adaptiveFvMesh::readDict() {                                  // adaptiveFvMesh.C:93-117
    const dictionary &refineDict = dict_.optionalSubDict("adaptiveFvMeshCoeffs");
    if (!refiner_) refiner_ = fvMeshRefiner::New(*this, refineDict);
    if (!error_) error_ = errorEstimator::New(mesh_, gpu_, refineDict);
    refiner_->readDict(refineDict);
    error_->read(refineDict);
}
void adaptiveFvMesh::meshChanged() {
    fvMeshLite &m = mesh_;
    std::vector<label> st, sz, ty, nb;
    for (auto &q : m.patches) { st.push_back(q.start); sz.push_back(q.size); ty.push_back(q.type); nb.push_back(q.nbrRank); }
    check(amr_mesh_set(gpu_, m.nCells, m.nInternalFaces, m.nBoundaryFaces, m.nPoints, m.owner.data(), m.neighbour.data(), (int)m.patches.size(),
                       st.data(), sz.data(), ty.data(), nb.data()), "amr_mesh_set");
    if (!m.pcOff.empty()) check(amr_mesh_set_points(gpu_, m.pcOff.data(), m.pcCells.data(), m.bndPoint.data()), "amr_mesh_set_points");
    if (!m.faceOff.empty()) check(amr_mesh_set_faces(gpu_, m.faceOff.data(), m.facePts.data()), "amr_mesh_set_faces");
    if (m.nEdges > 0) check(amr_mesh_set_edges(gpu_, m.nEdges, m.edgePts.data(), m.ecOff.data(), m.ecCells.data(), m.bndEdge.data()), "amr_mesh_set_edges");
    if (!m.Sf.empty()) check(amr_mesh_set_geometry(gpu_, m.weights.data(), m.Sf.data(), m.magSf.data(), m.deltaCoeffs.data(), m.V.data()), "amr_mesh_set_geometry");
    if (!m.cellLevel.empty())
        check(amr_levels_set(gpu_, m.cellLevel.data(), m.pointLevel.empty() ? nullptr : m.pointLevel.data(),
                             m.visibleCells.empty() ? nullptr : m.visibleCells.data(), m.splitParent.empty() ? nullptr : m.splitParent.data(),
                             (int64_t)m.splitParent.size()), "amr_levels_set");
    check(amr_protected_cells_set(gpu_, m.protectedCell.empty() ? nullptr : m.protectedCell.data()), "amr_protected_cells_set");
}
bool adaptiveFvMesh::update() {                                    // adaptiveFvMesh.C:369-376
    if (!refiner_ || !error_) readDict();
    bool changed = (refiner_->canRefine(true) || refiner_->canUnrefine(true)) && refine();
    return changed;
}
bool adaptiveFvMesh::refine() {                                    // adaptiveFvMesh.C:379-393
    dict_ = dictionary::parse(dictText_);     // "Re-read dictionary"
    readDict();
    error_->update();
    error_->protectPatches();
    return refiner_->refine(error_->error(), error_->maxRefinement());
}
void adaptiveFvMesh::updateMesh(const mapPolyMesh &map) {          // adaptiveFvMesh.C:293 -> fvMesh::mapFields
    // topoChanger has already replaced the addressing in mesh_; old sizes come from the map
    const int64_t nOld = (int64_t)map.reverseCellMap.size(), nNew = (int64_t)map.cellMap.size();
    for (auto &kv : mesh_.fields) {
        volField &f = kv.second;
        if ((int64_t)f.v.size() != nOld * f.nComp) continue;       // not a cell field of the old mesh
        scalarField nf((size_t)(nNew * f.nComp));
        check(amr_map_field(gpu_, nOld, nNew, map.cellMap.data(), f.nComp, f.v.data(), nf.data(), AMR_MAP_DIRECT, nullptr, nullptr),
              "fvMesh::mapFields");
        f.v.swap(nf);
    }
    meshChanged();
}

}  // namespace amr
