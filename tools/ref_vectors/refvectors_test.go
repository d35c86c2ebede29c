package playsnark

// Reference-vector dump for playsnark_b200 (tools/ref_vectors/README.md).
//
// Drop this file into a checkout of github.com/nikkolasg/playsnark (next to groth16.go; it must be
// in package playsnark because the structs' fields are unexported) and run, with the module's own
// pinned go.mod / go.sum:
//
//	PS_REF_OUT=/path/to/playsnark_b200/tests/golden go test -run TestDumpRefVectors -count=1 .
//
// It writes ref_readme.json and ref_chain16.json: every value the B200 backend and its oracle have
// to reproduce bit for bit, serialised with the reference's own MarshalBinary.  The trusted setups
// and the prover's (r, s) are random in the reference (Pick(random.New())); the dump records the
// toxic waste and the proof's retained (r, s), which is all the replay needs.

import (
	"encoding/hex"
	"encoding/json"
	"io/ioutil"
	"os"
	"path/filepath"
	"testing"

	"github.com/drand/kyber"
)

func hexScalar(s kyber.Scalar) string {
	b, err := s.MarshalBinary()
	if err != nil {
		panic(err)
	}
	return hex.EncodeToString(b)
}

func hexPoint(p kyber.Point) string {
	b, err := p.MarshalBinary()
	if err != nil {
		panic(err)
	}
	return hex.EncodeToString(b)
}

func hexPoints(ps []kyber.Point) []string {
	out := make([]string, 0, len(ps))
	for _, p := range ps {
		out = append(out, hexPoint(p))
	}
	return out
}

func hexPoly(p Poly) []string {
	out := make([]string, 0, len(p))
	for _, c := range p {
		out = append(out, hexScalar(c))
	}
	return out
}

func hexPolys(ps []Poly) [][]string {
	out := make([][]string, 0, len(ps))
	for _, p := range ps {
		out = append(out, hexPoly(p))
	}
	return out
}

func intVector(v Vector) []int {
	out := make([]int, 0, len(v))
	for _, x := range v {
		out = append(out, int(x))
	}
	return out
}

func intMatrix(m Matrix) [][]int {
	out := make([][]int, 0, len(m))
	for _, r := range m {
		out = append(out, intVector(r))
	}
	return out
}

// squaring chain x_{k+1} = x_k * x_k with n Mul gates; x0 = -1 keeps every value inside Value (int)
func chainCircuit(n int) (R1CS, Vector) {
	name := func(i int) string { return "x" + string(rune('A'+i/26)) + string(rune('a'+i%26)) }
	c := NewR1CS()
	c.NewInput(name(0))
	c.NewOutput(name(n))
	for i := 1; i < n; i++ {
		c.NewVar(name(i))
	}
	for i := 0; i < n; i++ {
		c.Mul(name(i), name(i), name(i+1))
	}
	sol := make(Vector, len(c.vars))
	sol[c.vars.IndexOf("const")] = 1
	v := Value(-1)
	for i := 0; i <= n; i++ {
		sol[c.vars.IndexOf(name(i))] = v
		v = v * v
	}
	return c, sol
}

func dumpCase(r1cs R1CS, sol Vector) map[string]interface{} {
	qap := ToQAP(r1cs)
	left, right, out := qap.computeAggregatePoly(sol)
	h := qap.Quotient(sol)
	diff := qap.nbVars - qap.nbIO

	tr := NewGroth16TrustedSetup(qap)
	proof := Groth16Prove(tr, qap, sol)
	if !Groth16Verify(tr, qap, proof, sol[:diff]) {
		panic("reference verifier rejects the reference proof")
	}
	g16 := map[string]interface{}{
		"toxic": map[string]string{"Alpha": hexScalar(tr.tw.Alpha), "Beta": hexScalar(tr.tw.Beta),
			"Delta": hexScalar(tr.tw.Delta), "X": hexScalar(tr.tw.X), "Gamma": hexScalar(tr.tw.Gamma)},
		"r": hexScalar(proof.tp.R), "s": hexScalar(proof.tp.S),
		"Alpha": hexPoint(tr.Alpha), "Beta": hexPoint(tr.Beta), "Delta": hexPoint(tr.Delta),
		"Beta2": hexPoint(tr.Beta2), "Delta2": hexPoint(tr.Delta2), "Gamma": hexPoint(tr.Gamma),
		"Xi": hexPoints(tr.Xi), "Xi2": hexPoints(tr.Xi2), "XiT": hexPoints(tr.XiT),
		"NioLP": hexPoints(tr.NioLP), "IoLP": hexPoints(tr.IoLP),
		"A": hexPoint(proof.A), "B": hexPoint(proof.B), "C": hexPoint(proof.C),
	}

	st := NewPHGR13TrustedSetup(qap)
	pp := PHGR13Prove(st.EK, qap, sol)
	if !PHGR13Verify(st.VK, qap, pp, sol[:diff]) {
		panic("reference verifier rejects the reference PHGR13 proof")
	}
	phgr := map[string]interface{}{
		"toxic": map[string]string{"s": hexScalar(st.t.s), "beta": hexScalar(st.t.beta), "rv": hexScalar(st.t.rv),
			"rw": hexScalar(st.t.rw), "ry": hexScalar(st.t.ry)},
		"ek": map[string][]string{"gsi": hexPoints(st.EK.gsi), "vs": hexPoints(st.EK.vs), "ws": hexPoints(st.EK.ws),
			"ys": hexPoints(st.EK.ys), "vas": hexPoints(st.EK.vas), "was": hexPoints(st.EK.was),
			"yas": hexPoints(st.EK.yas), "vbs": hexPoints(st.EK.vbs), "wbs": hexPoints(st.EK.wbs),
			"ybs": hexPoints(st.EK.ybs)},
		"proof": map[string]string{"hs": hexPoint(pp.hs), "vss": hexPoint(pp.vss), "wss": hexPoint(pp.wss),
			"yss": hexPoint(pp.yss), "vass": hexPoint(pp.vass), "wass": hexPoint(pp.wass),
			"yass": hexPoint(pp.yass), "gz": hexPoint(pp.gz)},
	}

	// wire-format anchors: small multiples of the generators and a few scalars
	var g1s, g2s, frs []string
	for _, k := range []int64{1, 2, 3, -1, 1 << 40} {
		e := NewElement().SetInt64(k)
		g1s = append(g1s, hexPoint(NewG1().Mul(e, nil)))
		g2s = append(g2s, hexPoint(NewG2().Mul(e, nil)))
		frs = append(frs, hexScalar(e))
	}

	return map[string]interface{}{
		"source":  "nikkolasg/playsnark, go test -run TestDumpRefVectors (tools/ref_vectors of playsnark_b200)",
		"witness": intVector(sol), "nb_vars": qap.nbVars, "nb_io": qap.nbIO, "nb_gates": qap.nbGates,
		"r1cs_left": intMatrix(r1cs.left), "r1cs_right": intMatrix(r1cs.right), "r1cs_out": intMatrix(r1cs.out),
		"left": hexPolys(qap.left), "right": hexPolys(qap.right), "out": hexPolys(qap.out), "z": hexPoly(qap.z),
		"a": hexPoly(left), "b": hexPoly(right), "c": hexPoly(out), "h": hexPoly(h),
		"groth16": g16, "phgr13": phgr,
		"wire":    map[string]interface{}{"k": []int64{1, 2, 3, -1, 1 << 40}, "g1": g1s, "g2": g2s, "fr": frs},
	}
}

func TestDumpRefVectors(t *testing.T) {
	dir := os.Getenv("PS_REF_OUT")
	if dir == "" {
		t.Skip("set PS_REF_OUT to the directory that receives ref_*.json")
	}
	write := func(name string, v interface{}) {
		b, err := json.MarshalIndent(v, "", " ")
		if err != nil {
			t.Fatal(err)
		}
		if err := ioutil.WriteFile(filepath.Join(dir, name), b, 0644); err != nil {
			t.Fatal(err)
		}
	}
	r1cs := createR1CS()
	write("ref_readme.json", dumpCase(r1cs, createWitness(r1cs)))
	c, sol := chainCircuit(16)
	write("ref_chain16.json", dumpCase(c, sol))
}
