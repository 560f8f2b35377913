"""``mbproj2.ModelNullPot`` work-alike: density, temperature and metallicity given directly."""


class ModelNullPot:
    def __init__(self, annuli, ne_cmpt, T_cmpt, Z_cmpt, NH_1022pcm2=None):
        self.annuli = annuli
        self.ne_cmpt = ne_cmpt
        self.T_cmpt = T_cmpt
        self.Z_cmpt = Z_cmpt
        self.NH_1022pcm2 = NH_1022pcm2

    def defPars(self):
        pars = self.ne_cmpt.defPars()
        pars.update(self.T_cmpt.defPars())
        pars.update(self.Z_cmpt.defPars())
        return pars

    def computeProfs(self, pars):
        return (self.ne_cmpt.computeProf(pars), self.T_cmpt.computeProf(pars),
                self.Z_cmpt.computeProf(pars))

    def prior(self, pars):
        return self.ne_cmpt.prior(pars) + self.T_cmpt.prior(pars) + self.Z_cmpt.prior(pars)
